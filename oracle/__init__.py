"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU (numpy) restatement of the surrDAMH hot path: Gaussian
random-walk proposal, Gaussian log-posterior, MH / delayed-acceptance MH accept
logic, polynomial (normalised Hermite) and RBF surrogates (apply + update), the
LHS initialiser and the analytic toy solver.  Every function cites the reference
file:line it follows.

Who may import this package: tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- and there only as the
checker / the timed CPU baseline, never as part of the product path.  The
product package (surrdamh_b200/) never imports oracle/ and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
section 4).  The restatement is therefore pinned against the reference itself, imported
unmodified from /root/reference in the build container through
oracle/ref_import.py (three environment shims, no edits), by
oracle/make_golden.py, which writes tests/golden/*.npz.  tests/test_oracle_*.py
check the restatement against those fixtures everywhere, and directly against the
live reference whenever /root/reference is present.
"""
