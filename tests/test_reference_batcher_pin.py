"""SURVEY.md 8 rows a12/a13 (the input tensor contract and the task batchers): batching.py against the REFERENCE'S OWN
loaders and minibatch iterators.

* where /root/reference exists (this container), tasks/qm9_task.py and tasks/ppi_task.py are executed unmodified under
  tests/tf1_shim (tf.placeholder as a feed_dict key, dpu_utils RichPath for local files -- nothing numerical is restated) on
  the 200 real QM9 validation molecules / a seeded PPI fold in the dgl layout, and every minibatch feed is compared with
  batching.py's: adjacency lists bit-exact INCLUDING edge order, graph ids, in-degrees, features, targets / labels, counts;
* everywhere (GPU box included) the same comparison runs against tests/golden/ref_batcher_feeds.npz, the reference feeds as
  written by tests/golden/make_batcher_fixtures.py;
* the two configurations the reference itself cannot run are pinned as such."""
import importlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.join(HERE, "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import batcher_cases as BC      # noqa: E402

batching = importlib.import_module("tf_gnn_samples_b200.batching")
FIXTURE = os.path.join(HERE, "golden", "ref_batcher_feeds.npz")
have_reference = pytest.mark.skipif(not os.path.isdir("/root/reference/tasks"), reason="the reference checkout is not on this box")


@pytest.fixture(scope="module")
def ppi_dir(tmp_path_factory):
    return BC.write_ppi_dir(str(tmp_path_factory.mktemp("ppi")), "test")


@pytest.fixture(scope="module")
def fixture():
    return np.load(FIXTURE)


@pytest.mark.parametrize("case", sorted(BC.QM9_CASES))
def test_qm9_feeds_equal_the_committed_reference_feeds(case, fixture):
    params, budget = BC.QM9_CASES[case]
    got, L = BC.repo_qm9_feeds(params, budget)
    assert L == int(fixture[case + "/num_edge_types"])
    BC.compare_feeds(got, BC.unpack_feeds(fixture, case), case)


@pytest.mark.parametrize("case", sorted(BC.PPI_CASES))
def test_ppi_feeds_equal_the_committed_reference_feeds(case, fixture, ppi_dir):
    params, budget = BC.PPI_CASES[case]
    got, L = BC.repo_ppi_feeds(params, budget, ppi_dir)
    assert L == int(fixture[case + "/num_edge_types"])
    BC.compare_feeds(got, BC.unpack_feeds(fixture, case), case)


@have_reference
@pytest.mark.parametrize("case", sorted(BC.QM9_CASES))
def test_qm9_feeds_equal_the_reference_loader_run_here(case, fixture):
    params, budget = BC.QM9_CASES[case]
    want, L = BC.reference_qm9_feeds(params, budget)
    got, L2 = BC.repo_qm9_feeds(params, budget)
    assert L == L2
    assert len(want) > 1 or budget >= 5000
    BC.compare_feeds(got, want, case)
    BC.compare_feeds(BC.unpack_feeds(fixture, case), want, case + " (fixture is current)")


@have_reference
@pytest.mark.parametrize("case", sorted(BC.PPI_CASES))
def test_ppi_feeds_equal_the_reference_loader_run_here(case, fixture, ppi_dir):
    params, budget = BC.PPI_CASES[case]
    want, L = BC.reference_ppi_feeds(params, budget, ppi_dir)
    got, L2 = BC.repo_ppi_feeds(params, budget, ppi_dir)
    assert L == L2
    BC.compare_feeds(got, want, case)
    BC.compare_feeds(BC.unpack_feeds(fixture, case), want, case + " (fixture is current)")


@have_reference
@pytest.mark.parametrize("case", sorted(BC.QM9_REFERENCE_RAISES))
def test_untied_qm9_cannot_run_in_the_reference(case):
    """qm9_task.py:139-145 appends to the list it enumerates -> IndexError on the first molecule.  batching.py builds what the
    loop evidently meant (forward types, then their reversals) instead of failing; stated here so the difference is on record."""
    params, budget = BC.QM9_REFERENCE_RAISES[case]
    with pytest.raises(IndexError):
        BC.reference_qm9_feeds(params, budget)
    feeds, L = BC.repo_qm9_feeds(params, budget)
    half = L // 2
    for f in feeds:
        for t in range(half):
            fwd, bwd = f["adjacency_e%d" % t], f["adjacency_e%d" % (half + t)]
            assert sorted(map(tuple, fwd[:, ::-1].tolist())) == list(map(tuple, bwd.tolist()))


@have_reference
def test_a_linkless_ppi_graph_breaks_the_reference_batcher_only(tmp_path):
    """A graph without links becomes np.array([]) of shape (0,) in ppi_task.py:152; packed next to a graph with links,
    np.concatenate (:247) raises.  batching.py keeps (0, 2) lists and packs it."""
    d = BC.write_ppi_dir(str(tmp_path), "test", linkless_graph=2)
    with pytest.raises(ValueError):
        BC.reference_ppi_feeds({}, 10 ** 6, d)
    feeds, L = BC.repo_ppi_feeds({}, 10 ** 6, d)
    assert len(feeds) == 1 and feeds[0]["num_graphs"] == 5


def test_minibatches_cover_every_graph_once_and_respect_the_budget():
    graphs = batching.make_qm9_like_graphs(300, seed=5)
    seen, budget = 0, 97
    for batch, first in batching.minibatches(graphs, budget):
        assert first == seen and batch.num_graphs >= 1 and batch.num_nodes < budget
        nxt = first + batch.num_graphs
        if nxt < len(graphs):            # the next graph is the one that did not fit (strict '<' of ppi_task.py:220)
            assert not (batch.num_nodes + graphs[nxt].node_features.shape[0] < budget)
        seen = nxt
    assert seen == len(graphs)


def test_minibatches_refuse_a_graph_that_can_never_fit():
    graphs = batching.make_qm9_like_graphs(3, seed=1)
    n = graphs[1].node_features.shape[0]
    with pytest.raises(ValueError, match="does not fit"):
        list(batching.minibatches(graphs, n))       # node_offset + n < n is false even for an empty batch


@have_reference
def test_the_full_qm9_validation_set_is_packed_like_the_reference():
    """BASELINE config 3's batch: all 10,000 validation molecules of data/qm9/valid.jsonl.gz through the reference's loader and
    batcher in ONE minibatch (V = 180,560, M = 554,026, L = 5) against batching.py -- every edge in the same position."""
    path = "/root/reference/data/qm9/valid.jsonl.gz"
    if not os.path.exists(path):
        pytest.skip("reference data not present")
    want, L = BC.reference_qm9_feeds({}, 10 ** 9, path=path)
    got, L2 = BC.repo_qm9_feeds({}, 10 ** 9, path=path)
    assert L == L2 == 5 and len(want) == 1
    assert (int(want[0]["num_graphs"]), int(want[0]["num_nodes"]), int(want[0]["num_edges"])) == (10000, 180560, 554026)
    BC.compare_feeds(got, want, "qm9 valid.jsonl.gz")
    # the structure-only archive the GPU box uses for config 3 (bench.py, test_reference_pin.py) rebuilds the same graph
    recs = batching.qm9_records_from_structure(os.path.join(HERE, "golden", "qm9_valid_structure.npz"))
    b, graph_nodes_list, _ = batching.qm9_batch(recs)
    assert np.array_equal(graph_nodes_list, want[0]["graph_nodes_list"])
    assert np.array_equal(b.type_to_num_incoming_edges, np.asarray(want[0]["type_to_num_incoming_edges"], np.float32))
    for i, a in enumerate(b.adjacency_lists):
        assert np.array_equal(a, want[0]["adjacency_e%d" % i]), "edge type %d" % i
