"""CPU oracle (test infrastructure only): see oracle/oracle.c.  Importable from tests/, smoke() and bench.py only."""
