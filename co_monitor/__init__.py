"""Import-path shim: lets code written against tfornal/co_monitor (`from co_monitor.geometry.simulation import
Simulation`, `from co_monitor.emissivity.reader import Emissivity`) run on the B200 implementation unchanged."""
