"""Stage 2: line-emissivity integration (host tables + ``Emissivity`` drop-in)."""
