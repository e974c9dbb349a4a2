"""Stand-in for the third-party `fastecdsa` package (AntonKueltz/fastecdsa).

TEST INFRASTRUCTURE ONLY.  The reference (wborgeaud/python-bulletproofs) delegates all
elliptic-curve arithmetic to `fastecdsa` (C + GMP), which is neither vendored, pinned,
nor installable in this image (no network, no gmp.h).  This package re-creates the small
API surface the reference touches in plain Python integers so the
UNMODIFIED reference can be imported from /root/reference by `oracle/gen_golden.py`.
It is never imported by the product package.
"""
