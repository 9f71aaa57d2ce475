"""-m gpu: a whole forward+backward step (NURBS evaluation, alignment, fused trace, per-target sum, loss, backward to the
control points) captured ONCE into a CUDA graph and replayed.  At the small sample counts of BASELINE.json configs 2 and 4
a step is bound by the host's launch rate (~25 launches + autograd bookkeeping: 0.7 ms for 0.15 ms of kernels,
tools/host_profile.py); every kernel of the path launches on torch's current stream through the C ABI, takes its
scratch from torch's allocator and does no device->host read in steady state (index checks and target counts are cached
per tensor version), so the step is capturable as it is."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _step(wl):
    total = wl.forward_local()
    loss = (total * total).mean()
    loss.backward()
    return total, loss


@pytest.mark.parametrize("kind", ["surface", "motor"])
def test_step_replays_as_a_cuda_graph_with_identical_results(kind):
    import bench

    dev = torch.device("cuda:0")
    wl = bench.Workload(dev, 12, 1, 0, kind=kind)
    for _ in range(3):                      # warm-up: caches (index ranges, grid detection, target count) are filled
        wl.param.grad = None
        _step(wl)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            wl.param.grad = None
            _step(wl)
    torch.cuda.current_stream().wait_stream(side)
    wl.param.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        total_g, loss_g = _step(wl)
    grad_g = wl.param.grad                  # static tensors of the captured step
    gen = torch.Generator(device=dev).manual_seed(5)
    for scale in (1e-4, 3e-4):
        with torch.no_grad():               # new parameters, in place (same storage the graph reads)
            wl.param.add_(scale * wl.param.abs().mean() * torch.randn(wl.param.shape, device=dev, generator=gen))
        graph.replay()
        got = (total_g.clone(), loss_g.clone(), grad_g.clone())
        wl.param.grad = None
        total_e, loss_e = _step(wl)
        assert got[0].sum() > 0
        assert torch.equal(got[0], total_e.detach()) and torch.equal(got[1], loss_e.detach())
        if kind == "surface":
            assert torch.equal(got[2], wl.param.grad)
        else:   # blocking: per-CTA double accumulators, practically (not provably) order-independent
            assert (got[2] - wl.param.grad).abs().max() <= 1e-6 * wl.param.grad.abs().max()
        wl.param.grad = grad_g              # hand the static gradient tensor back for the next replay
