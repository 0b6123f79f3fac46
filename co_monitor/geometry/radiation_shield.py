"""Drop-in imradiation_shield path: `co_monitor.geometry.radiation_shield` of tfornal/co_monitor, served by co_monitor_b200 (B200 implementation)."""
from co_monitor_b200.geometry.radiation_shield imradiation_shield *  # noqa: F401,F403
from co_monitor_b200.geometry imradiation_shield radiation_shield as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith('__')})
