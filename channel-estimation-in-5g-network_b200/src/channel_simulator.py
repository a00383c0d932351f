"""src.channel_simulator of the reference -> the B200 drop-in module of the same name (see src/__init__.py)."""
import channel_simulator as _impl
from channel_simulator import *  # noqa: F401,F403

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
