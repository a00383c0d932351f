"""src.utils of the reference -> the B200 drop-in module of the same name (see src/__init__.py)."""
import utils as _impl
from utils import *  # noqa: F401,F403

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
