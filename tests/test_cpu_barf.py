"""hold_b200.model.BarfSchedule against the weights the REFERENCE's BarfEmbedder holds along its schedule
(tests/golden/barf/barf_weights.pt, written by `python oracle/ref_harness.py golden_barf`)."""
import os

import torch


def test_barf_schedule_matches_reference_embedder():
    from hold_b200.model import BarfSchedule

    rec = torch.load(os.path.join(os.path.dirname(__file__), "golden", "barf", "barf_weights.pt"))
    for (start, end), ws in rec.items():
        s = BarfSchedule(6, 3, start, end)
        for it in range(max(ws) + 1):
            if it in ws:
                assert torch.equal(s.weights(), ws[it]), (start, end, it)
            s.step()
    assert s.weights().shape == (39,) and float(s.weights().min()) == 1.0     # past the end every frequency passes


def test_table_is_the_schedule():
    from hold_b200.model import BarfSchedule

    s = BarfSchedule(6, 3, 5, 25)
    t = s.table()
    assert t.shape == (25, 39) and s.alpha_iter == 0
    for it in range(25):
        assert torch.equal(t[it], s.weights())
        s.step()


def test_node_schedule_steps_in_place():
    """Node.start_barf / step_embedding (pure torch: run here on the CPU on a bare Node object): the weights buffer the kernels read
    keeps its address, follows the schedule step by step and saturates at the end."""
    import torch.nn as nn

    from hold_b200.model import BarfSchedule, LaplaceDensity, Node

    n = Node.__new__(Node)
    nn.Module.__init__(n)
    n.kind, n.density = "object", LaplaceDensity(0.1)
    n.start_barf(start=3, end=12)
    ref = BarfSchedule(6, 3, 3, 12)
    addr = n.barf_weights.data_ptr()
    for it in range(16):
        assert torch.equal(n.barf_weights, ref.weights()), it
        n.step_embedding()
        ref.step()
    assert n.barf_weights.data_ptr() == addr and float(n.barf_weights.min()) == 1.0
