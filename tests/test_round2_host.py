"""CPU: host-side logic added in round 2 (no GPU work)."""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_degree_balanced_cuts_match_the_torch_distributed_partition():
    from tf_gnn_samples_b200 import batching, degree_balanced_cuts
    from tf_gnn_samples_b200.partition import NodeRangePartition
    g = batching.make_typed_random_graph(500, 6000, (0.5, 0.3, 0.2), 4, seed=3)
    for world in (1, 2, 3, 8):
        cuts = degree_balanced_cuts(g.adjacency_lists, 500, world)
        assert cuts[0] == 0 and cuts[-1] == 500 and np.all(np.diff(cuts) >= 0)
        part = NodeRangePartition(g.adjacency_lists, g.type_to_node_to_num_incoming_edges, 500, 0, world)
        np.testing.assert_array_equal(cuts, part.cuts)
        indeg = g.type_to_node_to_num_incoming_edges.sum(axis=0) + 1
        loads = [indeg[cuts[r]:cuts[r + 1]].sum() for r in range(world)]
        assert max(loads) <= indeg.sum() / world + indeg.max() + 1          # balanced up to one node's weight


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_arms_share_one_config_dict():
    """The driver compares the `config` of our arm with the reference arm's: both print bench.CONFIG verbatim."""
    b = _load_bench()
    assert b.CONFIG["workload"] == b.WORKLOAD and b.CONFIG["M"] == 2 * b.NUM_LINKS + b.NUM_NODES
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": CONFIG') == 2
    assert json.loads(json.dumps(b.CONFIG)) == b.CONFIG
    assert b.algorithmic_bytes_per_layer(2245, 120245, 3, 256) == 129958012        # 130.0 MB per layer


def test_big_case_summary_bounds_follow_from_the_elementwise_tolerance():
    """ref_cases.compare_with_summary: an element-wise max-norm error eps implies every returned value <= eps."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_cases as RC
    rng = np.random.default_rng(0)
    out = rng.standard_normal((1000, 64))
    z = RC.summarize(out)
    eps = 3e-5
    noisy = out + eps * float(z["maxabs"]) * rng.uniform(-1, 1, size=out.shape)
    assert max(RC.compare_with_summary(noisy, z)) <= eps
    biased = out + eps * float(z["maxabs"])                                         # a systematic bias of eps is still within eps
    assert max(RC.compare_with_summary(biased, z)) <= eps * (1 + 1e-9)
    spoiled = out.copy(); spoiled[501, 7] += 1.0                                    # one bad element in an uncommitted row is seen
    assert max(RC.compare_with_summary(spoiled, z)) > 1e-4
