// Library-wide pieces of the C ABI: version, error string, geometry validation.
#include <stdarg.h>

#include "b2c_common.cuh"

namespace b2c {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_geom(const b2c_geom *g, bool allow_pitch) {
  B2C_REQUIRE(g->nsym >= 1 && g->nsym <= B2C_MAX_SYM, B2C_E_UNSUPPORTED, "nsym=%d outside [1,%d]", g->nsym,
              B2C_MAX_SYM);
  B2C_REQUIRE(g->ntx >= 1 && g->ntx <= B2C_MAX_ANT && g->nrx >= 1 && g->nrx <= B2C_MAX_ANT, B2C_E_UNSUPPORTED,
              "antenna counts %dx%d outside [1,%d]", g->ntx, g->nrx, B2C_MAX_ANT);
  // used bins = a block centred on DC with DC removed (src/channel_simulator.py:141-148): always an
  // odd count, symmetric in frequency; the slot kernel pairs bins -f / +f and draws them from one lane
  B2C_REQUIRE(g->nsc >= 1 && g->nsc <= 2 * B2C_RNG_LANES - 1 && (g->nsc & 1) == 1, B2C_E_UNSUPPORTED,
              "nsc=%d must be odd and <= %d", g->nsc, 2 * B2C_RNG_LANES - 1);
  B2C_REQUIRE(g->nsym * g->nsc <= 65535 * 4, B2C_E_UNSUPPORTED, "grid too large");
  B2C_REQUIRE(g->pitch == 0 || g->pitch == g->nsc || (allow_pitch && g->pitch > g->nsc), B2C_E_UNSUPPORTED,
              "row pitch %d: this entry point takes contiguous rows (pitch 0 or nsc=%d)", g->pitch, g->nsc);
  return B2C_OK;
}

}  // namespace b2c

extern "C" const char *b2c_last_error_string(void) { return b2c::g_err; }
extern "C" int b2c_abi_version(void) { return B2C_ABI_VERSION; }
