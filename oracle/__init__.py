"""CPU oracle (test infrastructure only; see oracle/oracle.h).  Never imported by the product."""
