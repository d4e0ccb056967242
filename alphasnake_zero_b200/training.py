"""Weight update for AlphaNNet.train (alpha_nnet.py:58-59, 78-106) -- OUTSIDE the self-play hot path.
Plain PyTorch autograd over the same weight dictionary: MSE + l2(1e-5) on every kernel, Adam with the reference's
piecewise-constant schedule (x0.25 after optimizer steps 20, 40, 60, 80; 0 after step 100: Keras `step <= boundary`), batch-norm in training mode."""
import numpy as np
import torch
import torch.nn.functional as F

from .utils.alpha_nnet import BN_EPS

L2 = 1e-5
BN_MOMENTUM = 0.99   # Keras default


def lr_at(step, lr):
    """alpha_nnet.py:79-84,92: Keras PiecewiseConstantDecay(boundaries=[20, 40, 60, 80, 100], values=[lr, lr/4, ..., lr/256, 0])
    evaluated at the optimizer's iteration counter (0 for the first update): values[0] while step <= 20, values[i] while
    boundaries[i-1] < step <= boundaries[i], 0 after step 100."""
    if step > 100:
        return 0.0
    return lr * 0.25 ** (max(step - 1, 0) // 20)


def fit(net, X, Y, epochs, batch_size, lr):
    dev = net.device
    w = net.weights
    def dev_tensor(a):      # device tensors (the engine's record gather) are used as they are; lists / arrays are uploaded once
        if torch.is_tensor(a):
            return a.to(dev, torch.float32)
        return torch.from_numpy(np.ascontiguousarray(np.array(a, dtype=np.float32))).to(dev)
    X = dev_tensor(X).permute(0, 3, 1, 2).contiguous()
    Y = dev_tensor(Y)
    P, bn_names, kernels = {}, [], []
    for k, v in w.items():
        if isinstance(v, dict):
            bn_names.append(k)
            P[k + ".gamma"] = torch.tensor(v["gamma"], device=dev, requires_grad=True)
            P[k + ".beta"] = torch.tensor(v["beta"], device=dev, requires_grad=True)
        elif isinstance(v, np.ndarray):
            P[k] = torch.tensor(v, device=dev, requires_grad=True)
            if v.ndim >= 2:
                kernels.append(k)
    stats = {k: (torch.tensor(w[k]["mean"], device=dev), torch.tensor(w[k]["var"], device=dev)) for k in bn_names}

    def bn(x, name):
        m, v = stats[name]
        y = F.batch_norm(x, m, v, P[name + ".gamma"], P[name + ".beta"], training=True, momentum=1 - BN_MOMENTUM, eps=BN_EPS)
        return y

    def conv(x, name):
        k = P[name].permute(3, 2, 0, 1)
        return F.conv2d(x, k, padding=k.shape[-1] // 2)

    def forward(x):
        h = F.relu(bn(conv(x, "conv0"), "bn0"))
        for b in range(4):
            sc = h
            h = F.relu(bn(conv(h, "res%d_conv0" % b), "res%d_bn0" % b))
            h = F.relu(bn(conv(h, "res%d_conv1" % b), "res%d_bn1" % b) + sc)
        h = F.relu(bn(conv(h, "head_conv"), "head_bn"))
        h = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)
        h = F.relu(h @ P["dense1_w"] + P["dense1_b"])
        return torch.tanh(h @ P["dense2_w"] + P["dense2_b"])

    opt = torch.optim.Adam(list(P.values()), lr=lr, eps=1e-7)
    step = 0
    n = X.shape[0]
    for _ in range(epochs):
        perm = torch.randperm(n, device=dev)
        for i in range(0, n, batch_size):
            idx = perm[i:i + batch_size]
            sched = lr_at(step, lr)
            for g in opt.param_groups:
                g["lr"] = sched
            loss = F.mse_loss(forward(X[idx]), Y[idx]) + L2 * sum((P[k] ** 2).sum() for k in kernels)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            step += 1
    out = {"side": w["side"]}
    for k, v in w.items():
        if isinstance(v, dict):
            out[k] = dict(gamma=P[k + ".gamma"].detach().cpu().numpy(), beta=P[k + ".beta"].detach().cpu().numpy(),
                          mean=stats[k][0].cpu().numpy(), var=stats[k][1].cpu().numpy())
        elif isinstance(v, np.ndarray):
            out[k] = P[k].detach().cpu().numpy()
    return out
