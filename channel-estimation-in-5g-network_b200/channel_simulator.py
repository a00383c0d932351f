"""Drop-in for the reference's src/channel_simulator.py, computed by libb2c on a B200.

Same names, arguments, return shapes/dtypes and random-draw order as the reference
(OFDMConfig, MIMOConfig, ChannelModel, OFDMSystem, PilotPattern, MIMOChannel,
simulate_transmission -- src/channel_simulator.py:17-421).  Randomness is drawn from the global
numpy.random stream in exactly the reference's order and injected into the kernels, so
`np.random.seed(s); simulate_transmission(...)` returns the reference's arrays (to fp32
accuracy; results are returned as complex128 NumPy arrays like the reference's).

Throughput work should not go through this per-slot surface: use `engine.SlotEngine.run` /
`dataset_generator.generate_batch`, which keep everything on the device and draw from Philox.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np
import torch

import _tables
from engine import SlotEngine

N_OSC = _tables.N_OSC


@dataclass
class OFDMConfig:
    """OFDM numerology (src/channel_simulator.py:17-24)."""
    fft_size: int = 1024
    cp_length: int = 72
    num_symbols: int = 14
    useful_subcarriers: int = 600
    subcarrier_spacing: float = 15000.0


@dataclass
class MIMOConfig:
    """Antenna counts (src/channel_simulator.py:27-31)."""
    num_tx: int = 2
    num_rx: int = 2


_ENGINES: Dict[tuple, SlotEngine] = {}


def _engine(ofdm: OFDMConfig, ntx: int, nrx: int) -> SlotEngine:
    """One engine (tables on the device) per geometry, reused across calls."""
    key = (ofdm.fft_size, ofdm.cp_length, ofdm.num_symbols, ofdm.useful_subcarriers,
           float(ofdm.subcarrier_spacing), ntx, nrx, torch.cuda.current_device() if torch.cuda.is_available() else -1)
    eng = _ENGINES.get(key)
    if eng is None:
        cfg = {"ofdm": {"fft_size": ofdm.fft_size, "cp_length": ofdm.cp_length, "num_symbols": ofdm.num_symbols,
                        "useful_subcarriers": ofdm.useful_subcarriers, "subcarrier_spacing": ofdm.subcarrier_spacing},
               "mimo": {"num_tx_antennas": ntx, "num_rx_antennas": nrx}}
        eng = _ENGINES[key] = SlotEngine(cfg)
    return eng


def _to_numpy(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().astype(np.complex128)


def _dev(a, dtype, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype)


class ChannelModel:
    """EPA / EVA / ETU tapped-delay-line fading (src/channel_simulator.py:34-127)."""

    CHANNEL_PROFILES = {name: {"delays": np.array(ns) * 1e-9, "powers": np.array(db)}
                        for name, (ns, db) in _tables.PDP.items()}

    def __init__(self, model_type: str, doppler_hz: float, carrier_freq: float, sampling_rate: float):
        self.model_type = model_type.upper()
        self.doppler_hz = doppler_hz
        self.carrier_freq = carrier_freq
        self.sampling_rate = sampling_rate
        self.delays, self.powers_db, self.powers_linear, self.delay_samples = _tables.path_tables(
            self.model_type, sampling_rate)      # KeyError on an unknown profile, as in the reference
        self.num_paths = len(self.delays)

    def _draw_jakes(self, num_tx: int, num_rx: int) -> np.ndarray:
        # reference order: path, tx, rx; rand(20) angles then rand(20) phases (:102-110)
        return np.random.rand(self.num_paths, num_tx, num_rx, 2, N_OSC)

    def generate_time_varying_channel(self, num_samples: int, num_tx: int, num_rx: int) -> np.ndarray:
        """(num_samples, num_rx, num_tx, max_delay+1) complex CIR, Jakes sum of sinusoids (:84-127)."""
        ju = self._draw_jakes(num_tx, num_rx)
        ofdm = OFDMConfig()
        ofdm.subcarrier_spacing = self.sampling_rate / ofdm.fft_size
        eng = _engine(ofdm, num_tx, num_rx)
        out = eng.tdl_full(self.model_type, self.doppler_hz, int(num_samples), num_tx, num_rx,
                           jakes_u=_dev(ju, torch.float32, eng.device))
        return _to_numpy(out)


class OFDMSystem:
    """Subcarrier map and OFDM modulate / demodulate (src/channel_simulator.py:130-203)."""

    def __init__(self, config: OFDMConfig):
        self.config = config
        self.sampling_rate = config.fft_size * config.subcarrier_spacing
        self.dc_idx = config.fft_size // 2
        self.used_indices = _tables.used_subcarriers(config.fft_size, config.useful_subcarriers)

    def modulate(self, symbols: np.ndarray) -> np.ndarray:
        """(num_symbols, used) -> (num_symbols, fft_size + cp_length) time samples with CP."""
        eng = _engine(self.config, 1, 1)
        return _to_numpy(eng.ofdm_modulate(_dev(symbols, torch.complex64, eng.device)))

    def demodulate(self, received_signal: np.ndarray) -> np.ndarray:
        """(num_symbols, fft_size + cp_length) -> (num_symbols, used)."""
        eng = _engine(self.config, 1, 1)
        return _to_numpy(eng.ofdm_demodulate(_dev(received_signal, torch.complex64, eng.device)))


class PilotPattern:
    """Random scattered pilots (src/channel_simulator.py:206-260): one global-RNG shuffle."""

    def __init__(self, num_subcarriers: int, num_symbols: int, pilot_density: float = 0.1):
        self.num_subcarriers = num_subcarriers
        self.num_symbols = num_symbols
        self.pilot_density = pilot_density
        total = num_subcarriers * num_symbols
        order = np.arange(total)
        np.random.shuffle(order)
        self.pilot_indices = np.sort(order[:int(total * pilot_density)])
        self.pilot_positions = np.unravel_index(self.pilot_indices, (num_symbols, num_subcarriers))
        self.pilot_mask = np.zeros((num_symbols, num_subcarriers), dtype=bool)
        self.pilot_mask[self.pilot_positions] = True

    def insert_pilots(self, data_symbols: np.ndarray, pilot_symbols: np.ndarray) -> np.ndarray:
        grid = np.zeros((self.num_symbols, self.num_subcarriers), dtype=complex)
        grid[self.pilot_mask] = pilot_symbols
        grid[~self.pilot_mask] = data_symbols
        return grid

    def extract_pilots(self, grid: np.ndarray) -> np.ndarray:
        return grid[self.pilot_mask]

    def get_pilot_positions(self) -> Tuple[np.ndarray, np.ndarray]:
        return self.pilot_positions


class MIMOChannel:
    """CFR generation and channel application (src/channel_simulator.py:263-345)."""

    def __init__(self, ofdm_config: OFDMConfig, mimo_config: MIMOConfig, channel_model: ChannelModel):
        self.ofdm_config = ofdm_config
        self.mimo_config = mimo_config
        self.channel_model = channel_model
        self.ofdm_system = OFDMSystem(ofdm_config)

    def _engine_for(self, num_symbols: int) -> SlotEngine:
        o = self.ofdm_config
        cfg = OFDMConfig(o.fft_size, o.cp_length, num_symbols, o.useful_subcarriers, o.subcarrier_spacing)
        return _engine(cfg, self.mimo_config.num_tx, self.mimo_config.num_rx)

    def _cfr(self, eng: SlotEngine, jakes_u: np.ndarray) -> torch.Tensor:
        cm = self.channel_model
        ju = _dev(jakes_u[None], torch.float32, eng.device)
        out = eng.run(1, eng.models.index(cm.model_type), cm.doppler_hz, 0.0, inject={"jakes_u": ju}, want=("H_true",))
        return out["H_true"][0]

    def generate_channel_frequency_response(self, num_symbols: int) -> np.ndarray:
        """(num_symbols, num_rx, num_tx, used) CFR sampled at each symbol start (:274-311)."""
        eng = self._engine_for(num_symbols)
        ju = self.channel_model._draw_jakes(self.mimo_config.num_tx, self.mimo_config.num_rx)
        return _to_numpy(self._cfr(eng, ju))

    def apply_channel(self, transmitted_symbols: np.ndarray, channel_response: np.ndarray, snr_db: float) -> np.ndarray:
        """y = Hx per resource element + AWGN at the slot's measured signal power (:313-345)."""
        nsym, nrx, ntx, nsc = channel_response.shape
        o = self.ofdm_config
        eng = _engine(OFDMConfig(o.fft_size, o.cp_length, nsym, o.useful_subcarriers, o.subcarrier_spacing), ntx, nrx)
        from _b2c import Geom
        g = Geom(nsym, nsc, ntx, nrx, o.fft_size, o.cp_length, eng.geom.symbol_period_s)
        z = np.random.randn(2, nsym, nrx, nsc)        # real block, then imaginary block (:342)
        noise = _dev(z[0] + 1j * z[1], torch.complex64, eng.device)[None]
        rx = eng.apply_channel(_dev(transmitted_symbols, torch.complex64, eng.device)[None],
                               _dev(channel_response, torch.complex64, eng.device)[None], float(snr_db),
                               noise=noise, geom=g)
        return _to_numpy(rx[0])


def simulate_transmission(config: Dict, channel_type: str = 'EPA', doppler_hz: float = 50,
                          snr_db: float = 10, pilot_density: float = 0.1) -> Dict:
    """One MIMO-OFDM slot (src/channel_simulator.py:348-421): same dict keys, same draw order."""
    o = config['ofdm']
    ofdm_cfg = OFDMConfig(o['fft_size'], o['cp_length'], o['num_symbols'], o['useful_subcarriers'],
                          o['subcarrier_spacing'])
    mimo_cfg = MIMOConfig(num_tx=config['mimo']['num_tx_antennas'], num_rx=config['mimo']['num_rx_antennas'])
    fs = ofdm_cfg.fft_size * ofdm_cfg.subcarrier_spacing
    model = ChannelModel(channel_type, doppler_hz, config['channel']['carrier_freq'], fs)
    eng = _engine(ofdm_cfg, mimo_cfg.num_tx, mimo_cfg.num_rx)
    nsym, nsc, ntx, nrx = eng.nsym, eng.nsc, eng.ntx, eng.nrx

    # draws, in the reference's order: shuffle, pilot phases, data phases, Jakes, noise
    pattern = PilotPattern(nsc, nsym, pilot_density)
    n_p = int(pattern.pilot_mask.sum())
    pilot_phase = np.random.uniform(0, 2 * np.pi, n_p)
    data_phase = np.random.uniform(0, 2 * np.pi, nsym * nsc - n_p)
    jakes_u = model._draw_jakes(ntx, nrx)
    z = np.random.randn(2, nsym, nrx, nsc)

    turns = np.empty((nsym, nsc))
    turns[pattern.pilot_mask] = pilot_phase / (2 * np.pi)
    turns[~pattern.pilot_mask] = data_phase / (2 * np.pi)
    inject = {"jakes_u": _dev(jakes_u[None], torch.float32, eng.device),
              "sym_turns": _dev(turns[None], torch.float32, eng.device),
              "noise": _dev((z[0] + 1j * z[1])[None], torch.complex64, eng.device)}
    out = eng.run(1, eng.models.index(model.model_type), doppler_hz, snr_db, inject=inject,
                  want=("H_true", "rx", "tx"))
    tx = _to_numpy(out["tx"][0])
    return {
        'tx_symbols': tx,
        'rx_symbols': _to_numpy(out["rx"][0]),
        'channel': _to_numpy(out["H_true"][0]),
        'pilot_pattern': pattern,
        'pilot_symbols': tx[:, 0, :][pattern.pilot_mask],
        'ofdm_config': ofdm_cfg,
        'mimo_config': mimo_cfg,
        'snr_db': snr_db,
    }
