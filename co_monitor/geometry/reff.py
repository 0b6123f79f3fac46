"""Drop-in import path: `co_monitor.geometry.reff` of tfornal/co_monitor, served by co_monitor_b200 (B200 implementation)."""
from co_monitor_b200.geometry.reff import *  # noqa: F401,F403
from co_monitor_b200.geometry import reff as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
