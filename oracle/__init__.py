"""CPU oracle for the biHomE hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``bihome_b200/`` (the product) may import this package.  The only
legitimate importers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``, and there only as
the checker / the timed CPU baseline, never as the thing shipped.

Layout
------
kornia050.py   restatement of the four kornia==0.5.0 functions the reference
               calls (third-party dependency pinned in reference
               ``requirements.txt:1``; NOT vendored in /root/reference).
ref_path.py    restatement of the reference's own orchestration of the path
               (``src/data/utils.py``, ``src/heads/PerceptualHead.py``,
               ``src/heads/ransac_utils.py``).
pairgen.py     restatement of the CPU pair generator (``src/data/transforms.py``)
               with explicit, injectable random draws.
ref_import.py  imports the UNMODIFIED reference modules from /root/reference
               (build container only) behind the kornia050 shim.
make_golden.py writes tests/golden/*.npz from ref_import (generator script).

Parity status: the reference ships no tests / golden vectors and its numerics
live in kornia 0.5.0, which cannot run on torch 2.11 (``torch.solve`` removed)
=> **parity is pinned against (i) the reference's own Python modules executed
here through the kornia050 restatement (tests/golden, made by make_golden.py),
(ii) OpenCV (getPerspectiveTransform / warpPerspective / findHomography) as an
independent implementation, (iii) fp64 autograd of the restatement.**  The
kornia layer itself is restated from its published 0.5.0 source, i.e. that
layer is "parity unpinned" in the strict sense (no reference-held vectors
exist for it); see DESIGN.md section 3.
"""
