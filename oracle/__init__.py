"""CPU oracle for the ODISE inference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (odise_b200/) never does and fails loudly when its CUDA library is missing.
Each function cites the reference file:line it restates.  Parity status per component is recorded in
DESIGN.md ("pinned" = checked here against the imported reference code or its own test vectors)."""
