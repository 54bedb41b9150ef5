"""CPU oracle package (test infrastructure only; see oracle/srvgg.py header)."""
