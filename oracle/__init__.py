"""CPU oracle for the rPPG signal path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the
timed CPU baseline.  Nothing under ``video-heart-rate_b200/`` imports it; the product
path fails loudly when the CUDA library is missing.

What is pinned and what is not (see DESIGN.md "Oracle"):

* ``oracle.roi`` / ``oracle.bpm`` restate reference functions that exist
  (``rppg_VIDEO.py``, ``rppg_LIVESTREAM.py``, ``analysis/utils/*.py``,
  ``analysis/measurement/green_avg.py``).  The reference ships no tests or golden
  vectors, so they are pinned by executing the reference's own function bodies
  verbatim (``oracle/ref_loader.py``, AST-extracted from ``/root/reference``) on
  seeded inputs; the resulting vectors are committed under ``tests/golden/`` with the
  script that made them (``tests/golden/make_golden.py``).
* ``oracle.evm`` (Gaussian pyramid, temporal ideal bandpass, amplify + collapse) and
  the polygon rasteriser in ``oracle.roi`` have NO reference implementation
  (SURVEY.md section 0.2/0.3): **parity unpinned** by the reference.  They are a frozen
  NumPy restatement of the published algorithm, cross-checked against the third-party
  arithmetic the reference's README points at (``cv2.pyrDown`` / ``cv2.pyrUp`` /
  ``np.fft``; cv2 4.13.0, numpy 2.3.5 in this image).
"""
