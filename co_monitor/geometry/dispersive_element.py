"""Drop-in import path: `co_monitor.geometry.dispersive_element` of tfornal/co_monitor, served by co_monitor_b200 (B200 implementation)."""
from co_monitor_b200.geometry.dispersive_element import *  # noqa: F401,F403
from co_monitor_b200.geometry import dispersive_element as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
