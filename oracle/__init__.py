"""CPU oracle for the relational message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``tf-gnn-samples_b200/``) imports from here.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` -- always as the checker / the CPU arm, never as the thing that
is shipped.

PARITY UNPINNED: the reference (microsoft/tf-gnn-samples) has no tests, golden
vectors or fixtures, and its arithmetic lives in TensorFlow 1.13 / dpu_utils
which cannot be installed in this image (no network, Python 3.12).  This oracle
is an op-for-op numpy restatement of ``gnns/*.py`` + ``utils/utils.py`` written
from the line-cited reference sources and the documented TF/Keras defaults
(SURVEY.md Appendix A).  The only reference-owned pin available is structural
(README.md:29 parameter count 699,257 -- see tests/test_oracle_known_answers.py).

Modules:
  ref_layers  numpy restatement (dtype-parametric: float64 "truth", float32
              "reference order") of the six sparse_*_layer functions.
  ref_torch   torch-CPU float32 restatement in reference op order (the timed
              stand-in for "the reference TF1 CPU path").
"""
