"""TEST INFRASTRUCTURE — independent cross-check of oracle/graph_oracle.py and the
multi-threaded CPU baseline ("port") that bench.py times.  Not product code.

Mirrors the *structure* of the reference TensorFlow graph with torch autograd:
  dM_dx = tf.gradients(model(Input), Input)[0][:, :dim]      TFModel.py:536-541
  loss  = w0*bCs + w1*iCs + w2*int2                          TFModel.py:643-666
  grad  = optimizer.compute_gradients(loss)                   TFModel.py:709
i.e. reverse-mode input gradient with create_graph=True followed by a second
reverse sweep w.r.t. the weights (double back-prop), which is what TF executes.
torch's multi-threaded CPU kernels are the closest available analogue of TF's
Eigen CPU executor (TF 1.10 cannot be installed here).
"""
import numpy as np
import torch

from .graph_oracle import layer_sizes, act_id, ACT_SIGMOID


def _split(theta, inpDim, layerWidth):
    Ws, bs, off = [], [], 0
    for i, o in layer_sizes(inpDim, layerWidth):
        Ws.append(theta[off:off + i * o].view(i, o)); off += i * o
        bs.append(theta[off:off + o]); off += o
    return Ws, bs


def _model(X, Ws, bs, act):
    a = X
    f = torch.sigmoid if act == ACT_SIGMOID else torch.tanh
    for l in range(len(Ws) - 1):
        a = f(a @ Ws[l] + bs[l])
    return a @ Ws[-1] + bs[-1]


def loss_and_grad(theta, feed, dim, inpDim, layerWidth, activation, timeDependent, lossOpt,
                  need_grad=True, dtype=torch.float64):
    act = act_id(activation)

    def c(v):   # float32 placeholder rounding, then arithmetic dtype
        return torch.as_tensor(np.asarray(v, dtype=np.float32)).to(dtype)

    th = c(theta).clone().requires_grad_(need_grad)
    Ws, bs = _split(th, inpDim, layerWidth)
    X = c(feed["Input"]).reshape(-1, inpDim).requires_grad_(True)
    nb, integNum = [int(v) for v in feed["intShape"]]
    P = nb * integNum
    u = _model(X, Ws, bs, act)                                    # [P,1]
    dM = torch.autograd.grad(u.sum(), X, create_graph=True)[0]     # tf.gradients(model(Input), Input)
    grad = dM[:, :dim]
    int1 = (grad * c(feed["gcoef"]).reshape(P, dim)).sum(-1, keepdim=True)
    if timeDependent:
        int1 = int1 - u * c(feed["dNt"]).reshape(P, 1)
    if lossOpt["isSource"]:
        int1 = int1 - c(feed["source"]).reshape(P, 1) * c(feed["N"]).reshape(P, 1)
    int1 = int1.reshape(nb, integNum)
    if lossOpt["integWflag"]:
        int1 = c(feed["integW"]).reshape(1, integNum) * int1
    R = int1.sum(-1, keepdim=True)
    int1 = R ** 2
    detJ = c(feed["detJ"])
    if bool(feed.get("detJvec", False)):
        detJ = detJ.reshape(nb, 1)
        int2 = (detJ * int1).sum()
    else:
        detJ = detJ.reshape(())
        int2 = detJ * int1.sum()
    lossVec = detJ * int1
    bX = c(feed["biInput"]).reshape(-1, inpDim)
    biVal = _model(bX, Ws, bs, act)
    biCs = float(np.float32(feed["biDimVal"])) * (biVal - c(feed["biLabel"]).reshape(-1, 1)) ** 2
    bDof = int(feed["bDof"])
    bCs = biCs[:bDof].mean()
    iCs = biCs[bDof:].mean() if timeDependent else torch.zeros((), dtype=dtype)
    w = c(feed["w"]).reshape(3)
    loss = w[0] * bCs + w[1] * iCs + w[2] * int2
    out = dict(loss=loss.item(), BCloss=bCs.item(), ICloss=iCs.item(), varLoss=int2.item(),
               lossVec=lossVec.detach().numpy().reshape(-1), R=R.detach().numpy().reshape(-1), grad=None)
    if need_grad:
        out["grad"] = torch.autograd.grad(loss, th)[0].detach().numpy()
    return out


class CpuStepper:
    """Pre-converted float32 tensors + one full step (loss + all weight gradients +
    TF-Adam) per call: the timed CPU baseline.  Conversion of the float64 feed to
    float32 (which the reference pays every sess.run, VarNetUtility.py:1044) is
    timed separately by bench.py."""

    def __init__(self, theta, feed, dim, inpDim, layerWidth, activation, timeDependent, lossOpt,
                 lr=1e-3, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        self.kw = dict(dim=dim, inpDim=inpDim, layerWidth=list(layerWidth), activation=activation,
                       timeDependent=timeDependent, lossOpt=lossOpt)
        self.act = act_id(activation)
        f32 = lambda v: torch.as_tensor(np.ascontiguousarray(np.asarray(v, dtype=np.float32)))
        self.nb, self.integNum = [int(v) for v in feed["intShape"]]
        P = self.nb * self.integNum
        self.X = f32(feed["Input"]).reshape(P, inpDim)
        self.gcoef = f32(feed["gcoef"]).reshape(P, dim)
        self.dNt = f32(feed["dNt"]).reshape(P, 1) if timeDependent else None
        self.srcN = (f32(feed["source"]).reshape(P, 1) * f32(feed["N"]).reshape(P, 1)) if lossOpt["isSource"] else None
        self.integW = f32(feed["integW"]).reshape(1, self.integNum) if lossOpt["integWflag"] else None
        self.detJ = float(np.float32(np.asarray(feed["detJ"]).reshape(-1)[0]))
        self.bX = f32(feed["biInput"]).reshape(-1, inpDim)
        self.bL = f32(feed["biLabel"]).reshape(-1, 1)
        self.bDof = int(feed["bDof"])
        self.biDimVal = float(np.float32(feed["biDimVal"]))
        self.w = [float(x) for x in np.asarray(feed["w"], dtype=np.float32).reshape(3)]
        self.theta = f32(theta).clone()
        self.m = torch.zeros_like(self.theta)
        self.v = torch.zeros_like(self.theta)
        self.t = 0
        self.lr = lr
        self.inpDim, self.dim, self.layerWidth = inpDim, dim, list(layerWidth)
        self.timeDependent = timeDependent

    def step(self):
        th = self.theta.clone().requires_grad_(True)
        Ws, bs = _split(th, self.inpDim, self.layerWidth)
        X = self.X.requires_grad_(True)
        u = _model(X, Ws, bs, self.act)
        dM = torch.autograd.grad(u.sum(), X, create_graph=True)[0]
        int1 = (dM[:, :self.dim] * self.gcoef).sum(-1, keepdim=True)
        if self.timeDependent:
            int1 = int1 - u * self.dNt
        if self.srcN is not None:
            int1 = int1 - self.srcN
        int1 = int1.reshape(self.nb, self.integNum)
        if self.integW is not None:
            int1 = self.integW * int1
        int2 = self.detJ * (int1.sum(-1) ** 2).sum()
        biCs = self.biDimVal * (_model(self.bX, Ws, bs, self.act) - self.bL) ** 2
        bCs = biCs[:self.bDof].mean()
        iCs = biCs[self.bDof:].mean() if self.timeDependent else 0.0
        loss = self.w[0] * bCs + self.w[1] * iCs + self.w[2] * int2
        g = torch.autograd.grad(loss, th)[0]
        self.t += 1
        b1, b2, eps = 0.9, 0.999, 1e-8
        self.m.mul_(b1).add_(g, alpha=1 - b1)
        self.v.mul_(b2).addcmul_(g, g, value=1 - b2)
        lr_t = self.lr * (1 - b2 ** self.t) ** 0.5 / (1 - b1 ** self.t)
        self.theta.addcdiv_(self.m, self.v.sqrt() + eps, value=-lr_t)
        return float(loss.detach())
