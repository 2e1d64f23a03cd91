"""CPU oracle for the multiview 2D->3D lifting hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or as the timed CPU baseline -- never as the thing shipped.  The
product path (``pose_unsupervised_b200``) never imports this package and has
no CPU fallback.

What it is: a numpy float64 restatement of the reference's algorithm for the
path named in BASELINE.json (reference checkout: LouisNUST/pose-unsupervised):

==========================  ====================================================
oracle module               follows (reference file:line)
==========================  ====================================================
``oracle.transforms``       lib/utils/transforms.py:67-135
``oracle.inference``        lib/core/inference.py:19-75
``oracle.cameras``          lib/multiviews/cameras.py:12-82
``oracle.pymvg_restated``   pymvg (un-pinned dependency, requirements.txt:13;
                            call sites lib/multiviews/triangulate.py:37-40,53,
                            147,210)
``oracle.triangulate``      lib/multiviews/triangulate.py:17-213
``oracle.body``             lib/multiviews/body.py:11-57
``oracle.pictorial``        lib/multiviews/pictorial.py:19-250 and
                            run/test/generate_pairwise_constraints.py:60-95
``oracle.epipolar``         lib/core/loss.py:101-133, run/test/test_fund_mtx.py:56-69
==========================  ====================================================

Parity pinning status
---------------------
* PINNED against the real reference code, executed in the build container by
  ``tests/golden/make_golden.py`` (imports ``/root/reference/lib``; the
  resulting vectors are committed under ``tests/golden/*.npz``):
  ``transforms``, ``inference``, ``cameras``, ``body``, ``pictorial``.
* PARITY UNPINNED at the pymvg boundary: ``pymvg`` is not vendored, not pinned
  to a version, not installed and not installable offline, and the reference
  holds no golden vectors for ``triangulate_poses`` / ``ransac`` /
  ``reproject_poses``.  ``oracle.pymvg_restated`` restates pymvg's published
  algorithm (Hartley & Zisserman linear triangulation with the 5-iteration
  OpenCV undistortion and the plumb-bob forward model) and is anchored by
  self-made known-answer tests (noise-free project->triangulate round trips,
  numpy SVD as the arithmetic reference, cv2.undistortPoints cross-check) and by
  an independent implementation of the same published algorithm: OpenCV's
  cv2.triangulatePoints agrees with the restated two-view find3d to <1e-6 mm.
* ``epipolar``: the formula is five lines of numpy in the reference's own
  evaluation script and is restated verbatim in meaning; the reference ships
  no fundamental-matrix pickle, so inputs are synthetic (F from cameras).
"""
