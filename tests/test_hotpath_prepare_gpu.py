"""HotPath(prepare_next=True): every step computes the frame-only work (identity losses, packed sources) of the NEXT input
set during its own backward and consumes what the previous step prepared.  The bookkeeping must never let a step consume
the preparation of frames that are no longer in its input set."""
import pytest
import torch

from _workload import baseline_config, head_state, make_host_batch

pytestmark = pytest.mark.gpu


def _hp(cfg, **kw):
    from sqlx.hotpath import HotPath
    hp = HotPath(cfg, **kw)
    hp.load_state_dict(head_state(cfg), strict=True)
    return hp


def _grads(hp):
    return [g.clone() for g in hp.param_grads()]


@pytest.mark.parametrize("use_graph", [False, True])
@pytest.mark.parametrize("num_slots,fork", [(1, "auto"), (2, "after_bwd_pred"), (3, "start"), (2, "after_pred_fwd")])
def test_prepare_next_matches_in_step_preparation(use_graph, num_slots, fork):
    cfg = baseline_config(2, B=2)
    batches = [make_host_batch(cfg, seed=100 + i) for i in range(4)]
    ref = _hp(cfg, use_graph=False, prepare_next=False)
    want = []
    for hb in batches:
        ref.load(hb, non_blocking=False)
        ref.step()
        torch.cuda.synchronize()
        want.append((float(ref.loss), _grads(ref)))
    hp = _hp(cfg, use_graph=use_graph, num_slots=num_slots, prepare_next=True, prepare_fork=fork)
    # a loader running ahead: batch i + 1 sits in the next set while batch i steps (when there is more than one set)
    hp.load(batches[0], non_blocking=False, slot=0)
    for i, hb in enumerate(batches):
        slot = i % num_slots
        if num_slots > 1 and i + 1 < len(batches):
            hp.load(batches[i + 1], non_blocking=False, slot=(i + 1) % num_slots)
        elif num_slots == 1:
            hp.load(hb, non_blocking=False, slot=0)          # new frames in the SAME set: must be re-prepared
        hp.step(slot)
        torch.cuda.synchronize()
        assert abs(float(hp.loss) - want[i][0]) < 1e-6, (i, float(hp.loss), want[i][0])
        for a, b in zip(_grads(hp), want[i][1]):
            assert float((a - b).abs().max()) <= 1e-6 * max(1.0, float(b.abs().max()))


def test_reload_after_preparation_is_detected():
    """step(0) prepares set 1 from frames A; loading frames B into set 1 afterwards must invalidate that preparation."""
    cfg = baseline_config(2, B=2)
    A, Bb = make_host_batch(cfg, seed=7), make_host_batch(cfg, seed=8)
    ref = _hp(cfg, use_graph=False, prepare_next=False)
    ref.load(Bb, non_blocking=False)
    ref.step()
    hp = _hp(cfg, use_graph=False, num_slots=2, prepare_next=True)
    hp.load(A, non_blocking=False, slot=0)
    hp.load(A, non_blocking=False, slot=1)
    hp.step(0)                       # prepares set 1 (frames A)
    hp.load(Bb, non_blocking=False, slot=1)
    hp.step(1)
    torch.cuda.synchronize()
    assert abs(float(hp.loss) - float(ref.loss)) < 1e-6
