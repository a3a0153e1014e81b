"""CPU oracle for the selective-pose glue path.  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
package, and only as the checker / the timed CPU reference; the product package
(``person-recognition-for-pose-estimation_b200``) never imports it.

Each function restates one reference function of SURVEY.md §8a and cites the reference
``file:line`` it follows (paths relative to ``/root/reference``; ``HF:`` =
``transformers/models/vitpose`` — third-party, pinned by the reference only as
``transformers>=4.48.1`` in ``requirements.txt:10``; 5.5.0 is installed in this image).

Pinning (how this oracle is tied to the reference):
* the reference ships **no** golden vectors / known-answer tests for this path, so
  ``oracle/gen_golden.py`` imports the reference's own functions from ``/root/reference`` in the
  build container, runs them on seeded inputs and commits the results under ``tests/golden/``;
  ``tests/test_oracle_golden.py`` checks every oracle function against those fixtures.
* pieces whose reference code is a third-party dependency that IS installed in this image
  (torchvision ``ops.nms``; HF ``VitPoseImageProcessor``) are additionally compared live against
  that dependency in the CPU tests.
* parity UNPINNED (no runnable reference anywhere): the gluoncv quarter-offset decode and
  ``get_affine_transform`` crop variant B (a10/a13 — gluoncv is neither vendored, pinned nor
  installed) and the similarity threshold gate (a8 — not in the reference).  Their restatements
  follow the published Simple-Baselines/HRNet formulas and are marked as such.
"""
from . import det, match, crop, pose, assoc  # noqa: F401
