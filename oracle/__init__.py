"""CPU oracle for the MISO-BF-MISO hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``misonet_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker or the reported CPU baseline.

Parity status: the reference (yuhogun0908/MISOnet @ 79b3190) ships no golden
vectors, known-answer tests or fixtures for this path, so
the oracle is pinned against *outputs of the reference itself*, produced in the
build container by importing ``/root/reference`` (see ``oracle/ref_import.py``)
and committed as ``tests/golden/*.npz`` by ``oracle/make_golden.py``.
"""
