"""Build helpers that let the reference's own Python layer (python-core) run on libdeciphon_b200.so."""
