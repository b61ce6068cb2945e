"""CPU oracle of the trace->proof path: test infrastructure only (see oracle/oracle.cc)."""
