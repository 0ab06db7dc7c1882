"""oracle/ -- CPU restatement of PhamClust's similarity-matrix hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under phamclust_b200/ may import this
package; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs do, and only as the checker or the timed CPU arm.

Parity status: gcs / jc / pocp / af and everything around the alignment are
pinned against the VERBATIM reference (run in the build container, golden
vectors under tests/golden/).  The alignment itself (third-party ``parasail``)
is a restatement of its published algorithm: PARITY UNPINNED at that boundary.
"""
