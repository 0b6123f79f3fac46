"""Drop-in import path: `co_monitor.geometry` of tfornal/co_monitor, served by co_monitor_b200 (B200 implementation)."""
