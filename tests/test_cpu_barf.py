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
