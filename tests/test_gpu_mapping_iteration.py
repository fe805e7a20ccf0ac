"""The bench's mapping iteration -- exactly as bench.py builds and runs it (gradient arena, CUDA-graph replay, two-stream
backward, fix_fine, colour decoder trained, bundle adjustment on 4 of 5 poses) -- against the oracle on the host cores
at the FULL size: 5 x 1000 rays x 48 samples on the Replica-room0 grids.  Every gradient of the iteration is compared:
three grids, every colour-decoder parameter (incl. the Fourier matrix B) and the four camera 7-vectors; then the same
iteration with the optimiser step inside."""
import pytest
import torch

import bench as B
from tests import helpers as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _box_independent_oracle():
    """The oracle runs on this box's CPU: pin the evaluation order of its Fourier argument (oracle/nice_oracle.py, EMBED_ORDER)
    so that the comparison does not depend on the host's BLAS."""
    from oracle import nice_oracle as NO
    with NO.fixed_embed_order():
        yield


DEV = torch.device("cuda", 0)


def _oracle_twin(w, with_optimizer):
    """The oracle arm started from the CUDA arm's decoders, grids and cameras (the frames are seeded identically)."""
    return B.OracleMapping(B.PIX_PER_KF, with_optimizer=with_optimizer,
                           sd={k: v.detach().cpu() for k, v in w.model.state_dict().items()},
                           grids={k: v.detach().cpu() for k, v in w.grids.items()}, cams=[c.detach().cpu() for c in w.cams])


@pytest.fixture(scope="module")
def pair():
    import pointnerf_slam_b200 as P
    from pointnerf_slam_b200 import engine as E
    w = B.build_mapping(DEV, 0, 1, B.PIX_PER_KF, "none", with_optimizer=False)
    om = _oracle_twin(w, with_optimizer=False)
    g = torch.Generator().manual_seed(99)
    idx = [torch.randint(B.H * B.W, (B.PIX_PER_KF,), generator=g) for _ in range(B.N_KEYFRAMES)]
    om.step(idx)
    yield w, om, idx
    E.GRAD_ARENA = None


def _check_grads(w, om, tag):
    msgs = []
    for k in ("grid_middle", "grid_fine", "grid_color"):
        # measured on a B200: 99.99 % of the elements within 3e-6 of the largest gradient element; a handful up to 2e-3
        # (a ReLU pre-activation within rounding of zero switches a unit on in one implementation and off in the other)
        gmax = float(om.c[k].grad.abs().max())
        msgs.append(T.assert_close_q(w.grids[k].grad, om.c[k].grad, rtol=0, atol=3e-6 * gmax, rtol_max=0, atol_max=5e-3 * gmax,
                                     q=0.999, what=f"{tag} grad {k}"))
        assert float((om.c[k].grad.abs().sum(1) > 0).float().mean()) < 0.5
    for name, p in w.model.named_parameters():
        if not name.startswith("color_decoder."):
            assert p.grad is None, name
            continue
        ref = om.sd[name].grad
        msgs.append(f"{tag} grad {name}: rel max err {T.rel_max(p.grad, ref):.2e}")
        assert T.rel_max(p.grad, ref) < 5e-4, (tag, name, T.rel_max(p.grad, ref))
    for i in range(1, B.N_KEYFRAMES):
        # The pose gradient sums 48,000 point gradients whose Fourier part cos(p.B) B^T (B ~ 25 N(0,1), untrained
        # decoders) is large and alternating: the sum is ill-conditioned, and the oracle's own float32 result moves in
        # the third digit with its summation order (measured: 6.5e-3 between the two).  The well-conditioned
        # quantities above (grids, decoder parameters) carry the parity claim; this bounds the pose gradient.
        msgs.append(f"{tag} grad camera {i}: rel max err {T.rel_max(w.cams[i].grad, om.cams[i].grad):.2e}")
        assert T.rel_max(w.cams[i].grad, om.cams[i].grad) < 3e-2, (tag, "camera", i)
    assert w.cams[0].grad is None
    return msgs


def test_full_size_mapping_gradients_eager_two_stream(pair):
    from pointnerf_slam_b200 import engine as E
    w, om, idx = pair
    assert E.PARALLEL_BACKWARD, "the bench default on one GPU"
    w.iteration.zero_grad()
    loss = w.iteration([i.to(DEV) for i in idx])
    assert loss.dtype == torch.float64
    del loss                      # (a failing test's traceback would keep the autograd graph alive into the capture below)
    print("\n".join(_check_grads(w, om, "eager")))


def test_full_size_mapping_gradients_graph_replay(pair):
    """The captured iteration (what bench.py times): pixel indices are read from static tensors, so the replay sees the
    oracle's draw; gradients after the replay are the oracle's."""
    import pointnerf_slam_b200 as P
    w, om, idx = pair
    static_idx = [i.to(DEV).clone() for i in idx]
    w.iteration.zero_grad()

    def body():
        for t in w.iteration.trained():
            t.grad = None
        return w.iteration(static_idx)
    gs = P.graphs.GraphedStep(body, generators=[w.gen], warmup=2)
    for t in static_idx:
        t.zero_()                      # prove the replay reads the tensors, not captured values
    gs()
    torch.cuda.synchronize()
    for t, i in zip(static_idx, idx):
        t.copy_(i)
    loss = gs()
    torch.cuda.synchronize()
    _check_grads(w, om, "graph")
    assert gs.launches <= 40, gs.launches
    gs.release()


def test_mapping_iteration_with_optimizer_tracks_reference(pair):
    """Three iterations with the frustum-masked Adam step inside (bench default): parameters follow the oracle's
    torch.optim.Adam-on-val[mask] loop (Mapper.py:509-518, 657-674)."""
    w2 = B.build_mapping(DEV, 0, 1, B.PIX_PER_KF, "none", with_optimizer=True)
    om = _oracle_twin(w2, with_optimizer=True)
    g = torch.Generator().manual_seed(7)
    before = {k: w2.grids[k].detach().clone() for k in w2.masks}
    for it in range(3):
        idx = [torch.randint(B.H * B.W, (B.PIX_PER_KF,), generator=g) for _ in range(B.N_KEYFRAMES)]
        om.step(idx)
        w2.iteration([i.to(DEV) for i in idx])
        w2.iteration.zero_grad()
    # Adam's first steps move an element by ~lr * sign(g) whatever |g| is, so an element whose gradient lies inside the
    # float32 noise of the two implementations (|g| < ~3e-6 of the largest) may step the other way: the comparison is
    # per element for the bulk and statistical for that tail; identical-gradient parity is in test_gpu_mapper.py
    for k, m in w2.masks.items():
        a, b = w2.grids[k].detach().cpu(), om.c[k].detach()
        diff = (a - b).abs()
        frac_off = float((diff > 1e-5).float().mean())
        print(f"{k}: {100 * frac_off:.4f}% of elements differ by more than 1e-5 after 3 Adam steps (max {float(diff.max()):.2e})")
        assert frac_off < 2e-3, k
        moved = (w2.grids[k].detach() != before[k]).any(1)[0]
        assert bool((moved <= m.bool()).all()), f"{k}: a voxel outside the frustum mask moved"
        assert int(moved.sum()) > 0
    for name, p in w2.model.named_parameters():
        if name.startswith("color_decoder."):
            d = (p.detach().cpu() - om.sd[name].detach()).abs()
            print(f"{name}: {100 * float((d > 1e-4).float().mean()):.3f}% of elements differ by more than 1e-4 (max {float(d.max()):.2e})")
            assert float((d > 1e-4).float().mean()) < 0.02, name
    for i in range(1, B.N_KEYFRAMES):
        assert T.rel_max(w2.cams[i], om.cams[i]) < 5e-3
    from pointnerf_slam_b200 import engine as E
    E.GRAD_ARENA = None
