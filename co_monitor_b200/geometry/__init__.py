"""Host-side geometry of the W7-X C/O monitor (setup classes + Stage-1 ``Simulation``)."""
