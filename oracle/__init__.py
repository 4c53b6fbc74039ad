"""ORACLE — test infrastructure, NOT product code.

CPU restatement of the reference's per-ray volumetric rendering hot path
(SURVEY.md §8a).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package, and only as the checker / the timed CPU baseline.  The product package
``cednerf_b200`` never imports it and has no CPU fallback.

Parity status
-------------
* In-tree reference code (``cednerf/encoder.py``, ``cednerf/render.py``,
  ``cednerf/utils.py``, ``cednerf/model.py``): PINNED.  ``tests/golden/make_golden.py``
  imports those files from ``/root/reference`` with their three missing
  third-party packages replaced by the restatements below and freezes
  input/output vectors under ``tests/golden/``; ``oracle.cednerf_ref`` is checked
  against them.
* Third-party arithmetic the reference only *calls* — nerfacc (marching,
  compositing), tiny-cuda-nn (hash grid, MLPs, Frequency/SH encodings) — and the
  Taichi kernels (need the Taichi compiler): PARITY UNPINNED.  None of the three
  is installed or installable here (no network, no version pins in the
  reference, no reference tests or golden vectors).  ``oracle.nerfacc_ref``,
  ``oracle.tcnn_ref`` and ``oracle.taichi_ref`` restate their published
  algorithms (SURVEY.md Appendix A/B and the in-tree Taichi source) and are the
  only pin that exists.

* ``torch.multinomial`` (the ISG / IST draw of ``datasets/dnerf_3d_video_IS.py:401-418``): PINNED to the real thing -
  torch is installed here, and ``oracle.dataset_ref.multinomial_without_replacement`` (top-k of weights / Exp(1) draws)
  equals ``torch.multinomial`` index for index under the same generator state
  (``tests/test_oracle_golden.py::test_multinomial_restatement_is_torch_multinomial``).
* ``torch_efficient_distloss`` (the ``-d`` regulariser, not vendored): its O(N) form is pinned to the O(N^2) definition it
  implements (``test_distortion_restatement_against_its_definition``).

Modules
-------
march_oracle.c   plain-C marcher (ray/aabb, boundary sort, traverse_grids), bit-exact contract
nerfacc_ref.py   ctypes wrapper + OccGridEstimator + volrend (torch, CPU)
tcnn_ref.py      hash grid, Frequency/SH encodings, fp16 fully-fused-MLP emulation
taichi_ref.py    the in-tree Taichi HashEncoder variants (3-D f16, 4-D key-frame)
cednerf_ref.py   DNGPradianceField, rendering, render_image, render_image_test, distortion
dataset_ref.py   the training branch of SubjectLoader.fetch_data (importance-sampled batches)
"""
