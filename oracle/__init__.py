"""CPU oracle for the SVD-Hybrid merge hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference algorithm
(mgradyn/SVD-Quantization-Task-Merging, ``src/svd_hybrid/*`` and the root
``quantization_utils.py``).  It exists to check the CUDA path; it is never the
thing that is shipped or measured as the product.

Who may import it
-----------------
* ``tests/``                          (as the checker)
* ``__graft_entry__.smoke()``         (as the checker)
* ``bench.py`` ``cpu_baseline`` leg and ``bench.py --impl reference``
  (timed as the *reference's* CPU path, never as ours)

Nothing under ``svd_quantization_task_merging_b200/`` imports this package.
The product path raises when the CUDA library is missing instead of falling
back to anything in here.

Layout
------
* ``svd_hybrid_ref.py``  torch-CPU restatement of the whole path.  The
  reference *is* torch-eager on CPU, so restating it with the same tensor
  runtime (``torch.linalg.svd`` -> LAPACK gesdd, ``torch.round`` half-to-even,
  separate fp32 mul/add) makes it bit-comparable with the reference on the
  same host.  Every function cites the reference file:line it follows.
* ``rtvq_ref.c``         plain-C restatement of the byte/integer pieces
  (tall-mask combination, asymmetric quantiser, multi-stage RTVQ, energy rank
  selection); built by ``oracle/build.py`` into ``oracle/liboracle_c.so``.
* ``cref.py``            ctypes binding for the C restatement.

Pinning
-------
Parity is PINNED: ``tests/golden/make_golden.py`` imported the real reference
from ``/root/reference`` in the build container and stored its inputs/outputs
as small fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this oracle against every one of them, and against the known-answer
vectors the reference's own tests hold (mask truth tables, the
``[1,2,3,4,5]``@4-bit case, the literal singular-value spectra).

Third-party arithmetic on the path that is not under /root/reference:
``torch.linalg.svd`` (torch 2.11.0, LAPACK gesdd through MKL; unpinned
upstream) and ``sklearn.cluster.KMeans`` (1.9.0 here; unpinned upstream).
Singular-vector signs and k-means label numbering are therefore not pinned by
the reference; the parity harness aligns signs and compares partitions.
"""
