"""tracking.TrackingIteration (src/Tracker.py:253-335) on the golden scene: the first iteration reproduces the reference's
tracking loss and camera gradient (goldens minted by oracle/pin_against_reference.py), its Adam step is torch.optim.Adam's,
three iterations follow a torch twin fed with the same gradients, and a CUDA-graph replay equals the eager iteration."""
import pytest
import torch

from tests import helpers as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _iteration(g, cam, with_opt=True):
    import pointnerf_slam_b200 as P
    from pointnerf_slam_b200.mapper import StageOptimizer
    from pointnerf_slam_b200.tracking import TrackingIteration
    model, grids, renderer = T.cuda_nice(g, DEV)
    H, W, fx, fy, cx, cy = [float(v) for v in g["intr"]]
    H0, H1, W0, W1 = [int(v) for v in g["crop"]]
    assert H0 == int(H) - H1 and W0 == int(W) - W1
    opt = None
    if with_opt:
        opt = StageOptimizer({}, [], [cam])
        opt.set_lrs([0.0, 0.0, 0.0, 0.0, 0.0, 1e-3])
    it = TrackingIteration(renderer, model, grids, g["depth_img"].to(DEV), g["color_img"].to(DEV), cam, int(H), int(W), fx, fy, cx, cy,
                           g["indices"].numel(), H0, W0, w_color=0.5, use_color=True, handle_dynamic=True, optimizer=opt)
    return it, model


def test_first_iteration_matches_the_reference_and_adam():
    g = T.load_nice()
    cam0 = g["cam"].clone()
    cam = cam0.to(DEV).requires_grad_(True)
    it, model = _iteration(g, cam)
    idx = g["indices"].to(DEV)
    loss = it(idx)
    torch.testing.assert_close(loss.cpu(), g["color/track/loss"].double(), rtol=2e-3, atol=1e-3)
    assert T.rel_max(cam.grad, g["color/track/grad_cam"]) < 5e-3
    assert all(p.grad is None for p in model.parameters())          # the map is frozen
    # first Adam step: cam - lr * g / (|g| + eps)
    twin = cam0.clone().requires_grad_(True)
    opt = torch.optim.Adam([twin], lr=1e-3)
    twin.grad = cam.grad.detach().cpu().clone()
    opt.step()
    assert (cam.detach().cpu() - twin.detach()).abs().max() < 1e-7
    # two more iterations: the twin takes the CUDA gradients, the parameters must follow torch.optim.Adam
    for _ in range(2):
        it(idx)
        twin.grad = cam.grad.detach().cpu().clone()
        opt.step()
    assert (cam.detach().cpu() - twin.detach()).abs().max() < 2e-7
    assert (cam.detach().cpu() - cam0).abs().max() > 1e-3             # and it did move


def test_forward_loss_backward_equals_the_direct_path():
    g = T.load_nice()
    idx = g["indices"].to(DEV)
    cam_a = g["cam"].to(DEV).requires_grad_(True)
    it_a, _ = _iteration(g, cam_a, with_opt=False)
    la = it_a(idx)
    cam_b = g["cam"].to(DEV).requires_grad_(True)
    it_b, _ = _iteration(g, cam_b, with_opt=False)
    lb = it_b.forward_loss(idx)
    lb.backward()
    assert torch.equal(la, lb.detach())
    torch.testing.assert_close(cam_a.grad, cam_b.grad, rtol=1e-5, atol=1e-6)     # atomics order only


def test_graph_replay_equals_eager():
    import pointnerf_slam_b200 as P
    g = T.load_nice()
    idx = g["indices"].to(DEV)
    cam_e = g["cam"].to(DEV).requires_grad_(True)
    it_e, _ = _iteration(g, cam_e)
    for _ in range(3):
        it_e(idx)
    cam_g = g["cam"].to(DEV).requires_grad_(True)
    it_g, _ = _iteration(g, cam_g)
    step = P.graphs.GraphedStep(lambda: it_g(idx), warmup=1)
    # the warm-up pass ran the iteration for real: restore the start state, then replay three times
    cam_g.data.copy_(g["cam"].to(DEV))
    it_g.optimizer.step_count.zero_(); [m.zero_() for m in it_g.optimizer.exp_avg]; [v.zero_() for v in it_g.optimizer.exp_avg_sq]
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    assert (cam_g.detach() - cam_e.detach()).abs().max() < 1e-6
    step.release()
