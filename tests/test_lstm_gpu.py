"""Persistent LSTM kernel (csrc/lstm_kernels.cu) against the plain PyTorch loop and ``torch.nn.LSTM`` in fp32."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T,I,H", [(4, 80, 8, 256), (3, 17, 46, 64), (9, 12, 32, 128), (4, 80, 256, 256)])
@pytest.mark.parametrize("with_state", [False, True])
def test_lstm_layer_matches_reference(B, T, I, H, with_state):
    from msrflute_b200.ops import lstm_ops
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    k = 1.0 / H ** 0.5
    x = torch.randn(B, T, I, device="cuda", requires_grad=True)
    ws = [torch.empty(4 * H, I, device="cuda").uniform_(-k, k).requires_grad_(True),
          torch.empty(4 * H, H, device="cuda").uniform_(-k, k).requires_grad_(True),
          torch.empty(4 * H, device="cuda").uniform_(-k, k).requires_grad_(True),
          torch.empty(4 * H, device="cuda").uniform_(-k, k).requires_grad_(True)]
    h0 = torch.randn(B, H, device="cuda", requires_grad=True) if with_state else None
    c0 = torch.randn(B, H, device="cuda", requires_grad=True) if with_state else None
    dy = torch.randn(B, T, H, device="cuda")
    dh, dc = torch.randn(B, H, device="cuda"), torch.randn(B, H, device="cuda")

    def run(fn):
        for t in [x] + ws + ([h0, c0] if with_state else []):
            t.grad = None
        hs, (hT, cT) = fn(x, *ws, h0, c0)
        ((hs * dy).sum() + (hT * dh).sum() + (cT * dc).sum()).backward()
        return [hs.detach(), hT.detach(), cT.detach(), x.grad.clone()] + [w.grad.clone() for w in ws] + \
               ([h0.grad.clone(), c0.grad.clone()] if with_state else [])

    got = run(lstm_ops.lstm_layer)
    want = run(lstm_ops.lstm_layer_reference)
    names = ["hs", "hT", "cT", "dx", "dw_ih", "dw_hh", "db_ih", "db_hh", "dh0", "dc0"]
    for n, a, b in zip(names, got, want):
        err = float((a - b).norm() / (b.norm() + 1e-12))
        assert err < 2e-4, (n, err)


def test_lstm_module_matches_nn_lstm_and_loads_its_state():
    from msrflute_b200.ops.lstm_ops import LSTM
    torch.manual_seed(1)
    ref = torch.nn.LSTM(8, 256, num_layers=2, batch_first=True).cuda()
    ours = LSTM(8, 256, num_layers=2).cuda()
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(4, 80, 8, device="cuda")
    y1, (h1, c1) = ref(x)
    y2, (h2, c2) = ours(x)
    assert torch.allclose(y1, y2, atol=2e-5, rtol=1e-4)
    assert torch.allclose(h1, h2, atol=1e-4, rtol=1e-3) and torch.allclose(c1, c2, atol=1e-4, rtol=1e-3)
    y1.sum().backward()
    y2.sum().backward()
    for (n, p), (_, q) in zip(ref.named_parameters(), ours.named_parameters()):
        assert torch.allclose(p.grad, q.grad, atol=2e-3, rtol=2e-3), n
