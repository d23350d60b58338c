"""GPU: the DGL-free conv layers / models and the training loop (f-2) against a plain
PyTorch fp32 restatement of the same formulas (torch.topk + dense scatter + torch.sparse mm)."""
import pytest
import torch
import torch.nn as nn

from helpers import make_problem
from synth_graphs import symmetrize

pytestmark = pytest.mark.gpu


def _graph(n=400, e=6000, seed=3):
    from maxk_models_integrated import CSRGraph
    g = symmetrize(make_problem(n, e, 8, seed=seed)["graph"])
    gc = CSRGraph.from_dict({k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in g.items()}, "toy")
    a = torch.sparse_csr_tensor(g["indptr"].long(), g["indices"].long(), g["values"], size=(n, n)).cuda()
    return gc, a


def _ref_aggregate(a, deg, x, k):
    """(A @ (x * topk_mask)) / deg in plain torch (what the reference computes via DGL / cuSPARSE)."""
    v, i = torch.topk(x, k, dim=1)
    xs = torch.zeros_like(x).scatter(1, i, v)
    return (a @ xs) / deg.unsqueeze(-1)


@pytest.mark.parametrize("which", ["sage", "gcn", "gin"])
def test_model_forward_and_gradients_match_torch_reference(which):
    from maxk_models_integrated import MaxKGCN, MaxKGIN, MaxKSAGE
    torch.manual_seed(0)
    gc, a = _graph()
    n, k, hid = gc.num_nodes(), 16, 256
    cls = {"sage": MaxKSAGE, "gcn": MaxKGCN, "gin": MaxKGIN}[which]
    model = cls(32, hid, 2, 7, maxk=k, feat_drop=0.0, norm=True, graph_name="toy").cuda()
    x = torch.randn(n, 32, device="cuda")
    out = model(gc, x)
    loss = out.square().mean()
    loss.backward()
    grads = {name: p.grad.clone() for name, p in model.named_parameters() if p.grad is not None}

    # same weights, plain torch ops
    deg = gc.degrees

    def ref_forward():
        if which == "sage":
            h = model.lin_in(x)
            for layer in model.layers:
                agg = _ref_aggregate(a, deg, h, k)
                hs = h * torch.zeros_like(h).scatter(1, torch.topk(h, k, dim=1)[1], 1.0)
                h = layer.norm(layer.fc_self(hs) + layer.fc_neigh(agg))
            return model.lin_out(h)
        h = model.lin_in(x).relu()
        for i in range(model.num_layers):
            h = model.linlayers[i](h)
            agg = _ref_aggregate(a, deg, h, k)
            hs = h * torch.zeros_like(h).scatter(1, torch.topk(h, k, dim=1)[1], 1.0)
            if which == "gcn":
                h = agg * torch.pow(deg, -0.5).unsqueeze(-1)
            else:
                h = (1 + model.convlayers[i].eps) * hs + agg
            h = model.normlayers[i](h)
        return model.lin_out(h)

    model.zero_grad()
    ref = ref_forward()
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-5)
    ref.square().mean().backward()
    for name, p in model.named_parameters():
        if p.grad is not None:
            torch.testing.assert_close(grads[name], p.grad, rtol=2e-3, atol=1e-5, msg=lambda m, n=name: n + ": " + m)


def test_training_loop_reduces_loss_and_reports_timing():
    from maxk_gnn_training import synthetic_task, train
    from maxk_models_integrated import MaxKSAGE
    torch.manual_seed(1)
    graph, x, y, masks = synthetic_task("flickr", 0.02, 64, 5, torch.device("cuda"))
    model = MaxKSAGE(64, 256, 2, 5, maxk=32, feat_drop=0.1, graph_name="flickr").cuda()
    rep = train(graph, x, y, masks, model, epochs=25, warmup_epochs=5, log=None)
    assert rep["epochs_measured"] == 20 and rep["avg_forward_ms"] > 0 and rep["avg_backward_ms"] > 0
    assert rep["losses"][-1] < rep["losses"][0]


def test_layers_validate_inputs():
    from maxk_models_integrated import MaxKGINConv, MaxKGraphConv, MaxKSAGEConv
    gc, _ = _graph(100, 800)
    x = torch.randn(100, 256, device="cuda")
    with pytest.raises(RuntimeError):
        MaxKSAGEConv(256, 256).cuda()(gc, x)
    with pytest.raises(ValueError):
        MaxKGraphConv(256, 256, norm="bogus")
    with pytest.raises(KeyError):
        MaxKGINConv(aggregator_type="max")
    with pytest.raises(ValueError):
        MaxKSAGEConv(256, 256, aggregator_type="pool")
    assert isinstance(MaxKSAGEConv(256, 256, norm=nn.LayerNorm(256)).norm, nn.LayerNorm)
