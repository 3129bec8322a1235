"""CPU oracle for the relational message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``tf-gnn-samples_b200/``) imports from here.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` -- always as the checker / the CPU arm, never as the thing that
is shipped.

PINNED AGAINST THE REFERENCE'S OWN CODE (round 2): the reference (microsoft/tf-gnn-samples) has no tests, golden
vectors or fixtures, and its arithmetic lives in TensorFlow 1.13 / dpu_utils which cannot be installed in this image (no
network, Python 3.12) -- but ``gnns/*.py`` and ``utils/utils.py`` are plain Python over ~25 ``tf.*`` calls, so
``tests/golden/make_ref_fixtures.py`` imports them UNMODIFIED with a numpy-backed ``tensorflow`` / ``dpu_utils`` stand-in
(``tests/tf1_shim``), feeds them seeded inputs and weights, and commits what they return as ``tests/golden/ref_*.npz``.
``tests/test_reference_pin.py`` holds this oracle to those files at 1e-12 (float64) for 17 small cases and for BASELINE.json
configs 2-5 at full size.  Remaining assumption: the TF 1.13 / Keras / dpu_utils kernel semantics restated by the stand-in
(SURVEY.md Appendix A); ``tests/golden/make_tf1_fixtures.py`` checks them against a real TensorFlow 1.13 where one exists.
The structural known answer (README.md:29 parameter count 699,257) stays in tests/test_oracle_known_answers.py.

Modules:
  ref_layers  numpy restatement (dtype-parametric: float64 "truth", float32
              "reference order") of the six sparse_*_layer functions.
  ref_torch   torch-CPU float32 restatement in reference op order (the timed
              stand-in for "the reference TF1 CPU path"; pinned by tests/test_cpu_baseline_pin.py).
  ref_model   whole-model forward + task heads (pinned against the reference's own scaffold:
              tests/test_reference_model_pin.py).
"""
