"""ORACLE — test infrastructure only (PARITY UNPINNED: the reference mount ships no source, no golden
vectors and no tests; see oracle/networks.py header and DESIGN.md).  Never imported by the product
package; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it."""
