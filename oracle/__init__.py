"""CPU oracle for the ensemble Kalman update path of agarbuno/ces.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ces_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker or as the timed CPU baseline, never as the product path.

Modules
-------
eks_oracle       numpy restatement of ``sampling.eks_update*`` /
                 ``timestep_method`` / metrics (ces/calibrate.py:243-529).
                 Parity PINNED: checked against the real reference (run here,
                 tab-expanded in memory) and against tests/golden/*.npz that
                 were generated from it (tests/golden/make_golden.py).
forward_oracle   numpy restatement of the ``ces.utils`` map-type forward
                 models (ces/utils.py:5-122).  Parity PINNED the same way.
darcy_oracle     scipy restatement of ces/darcy.py + utilities/mfiles/*.m.
                 PARITY UNPINNED: the reference needs a MATLAB engine that is
                 not available and stores no Darcy output anywhere.
reference_loader loads the real reference from /root/reference when present
                 (never at run time on the GPU box).
"""
