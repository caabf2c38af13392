"""CPU oracle for the HardNet hot path — TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package. The product path (hardnetnas_b200/) never does: it fails loudly when the CUDA extension is
missing.

The reference is pure Python/PyTorch, so the restatement uses plain torch CPU functional ops in fp32
(the reference's own arithmetic library; oneDNN / MKL underneath). Every function cites the reference
file:line it follows. PARITY IS PINNED: tests/golden/*.npz hold outputs of the reference's own code,
imported unmodified from /root/reference by oracle/make_golden.py (committed), and
tests/test_oracle_golden.py checks every oracle function against them.
"""
