"""TEST-ONLY stand-in for varnet_b200._capi.Engine backed by the FP64 oracle.

Lets the CPU suite exercise the `TFNN` shim's protocol logic (feed caching, fetch semantics, tower
sums, checkpointing) — including under the reference's own unmodified `VarNet.train` — in a container
without a GPU.  It is injected by monkeypatching inside tests only; the product never imports it
(the real Engine raises when no CUDA device is present)."""
import numpy as np

from oracle import graph_oracle as go


class FakeCfg:
    device = 0


class FakeEngine:
    instances = []

    def __init__(self, dim, inpDim, layerWidth, activation="sigmoid", timeDependent=True, isSource=False,
                 integWflag=False, optimizer="adam", device=0):
        self.dim, self.inpDim, self.layerWidth = dim, inpDim, list(layerWidth)
        self.kw = dict(dim=dim, inpDim=inpDim, layerWidth=self.layerWidth, activation=activation,
                       timeDependent=timeDependent, lossOpt=dict(isSource=isSource, integWflag=integWflag))
        self.optimizer = optimizer
        self.cfg = FakeCfg(); self.cfg.device = device
        self.nparam = go.param_count(inpDim, self.layerWidth)
        self.theta = np.zeros(self.nparam); self.m = np.zeros(self.nparam); self.v = np.zeros(self.nparam)
        self.step = 0
        self.feed = {"w": np.ones(3)}          # the real engine also starts with unit loss weights
        self.calls = dict(upload_points=0, upload_bic=0, loss=0, loss_grad=0, loss_grad_fed=0, step=0, eval=0)
        self.gbuf = np.zeros(self.nparam + 4)
        FakeEngine.instances.append(self)

    def close(self):
        pass

    def set_params(self, theta):
        self.theta = np.asarray(theta, dtype=np.float32).astype(np.float64).copy()
        self.m[:] = 0; self.v[:] = 1.0 if self.optimizer.startswith("rms") else 0.0; self.step = 0

    def get_params(self):
        return self.theta.astype(np.float32)

    def get_optimizer_state(self):
        return self.m.astype(np.float32), self.v.astype(np.float32), self.step

    def set_optimizer_state(self, m, v, step):
        self.m, self.v, self.step = np.array(m, dtype=np.float64), np.array(v, dtype=np.float64), int(step)

    supports_table_views = True

    def upload_points(self, Input, gcoef, source, N, dNt, intShape, integW, detJ, detJvec=False, dtype=None):
        self.upload_table(Input, gcoef, source, N, dNt, intShape, integW, detJ, detJvec, nx=self.inpDim)
        self.set_extra_inputs(None)

    # device-resident tables / mini-batches: same call protocol as the real engine, materialised for the oracle
    def select_table(self, slot):
        self.slot = slot
        self.tables = getattr(self, "tables", {})
        if slot in self.tables:
            self.batch = None
            self._refresh()

    def upload_table(self, Input, gcoef, source, N, dNt, intShape, integW, detJ, detJvec=False, dtype=None, nx=None):
        self.tables = getattr(self, "tables", {})
        self.slot = getattr(self, "slot", 0)
        arr = lambda a: None if a is None or np.asarray(a).dtype == object else np.array(a, dtype=np.float64)
        self.tables[self.slot] = dict(Input=arr(Input), gcoef=arr(gcoef), source=arr(source), N=arr(N), dNt=arr(dNt),
                                      nbTab=int(intShape[0]), integNum=int(intShape[1]), integW=integW, detJ=detJ,
                                      detJvec=detJvec)
        self.batch = None
        self.calls["upload_points"] += 1
        self._refresh()

    def set_batch(self, tf_index):
        self.batch = None if tf_index is None else np.array(tf_index, dtype=np.int64).ravel()
        self.calls["set_batch"] = self.calls.get("set_batch", 0) + 1
        self._refresh()

    def set_extra_inputs(self, vals):
        self.extra = None if vals is None or np.size(vals) == 0 else np.asarray(vals, dtype=np.float32).astype(np.float64).reshape(1, -1)
        if getattr(self, "tables", None) and getattr(self, "slot", 0) in self.tables:
            self._refresh()

    def _refresh(self):
        t = self.tables[self.slot]
        q = t["integNum"]
        tf = np.arange(t["nbTab"]) if getattr(self, "batch", None) is None else self.batch
        rows = (tf.reshape(-1, 1) * q + np.arange(q)).ravel()
        take = lambda a: None if a is None else a.reshape(t["nbTab"] * q, -1)[rows]
        X = take(t["Input"])
        extra = getattr(self, "extra", None)
        if extra is not None:
            X = np.hstack([X, np.tile(extra, [len(X), 1])])
        detJ = t["detJ"]
        if t["detJvec"]:
            detJ = np.asarray(detJ, dtype=np.float64).reshape(-1, 1)[tf]
        self.feed.update(Input=X, gcoef=take(t["gcoef"]), source=take(t["source"]), N=take(t["N"]),
                         dNt=take(t["dNt"]) if t["dNt"] is not None else [[None]], intShape=[len(tf), q],
                         integW=t["integW"], detJ=detJ, detJvec=t["detJvec"])
        self.nb = len(tf)

    def upload_bic(self, biInput, biLabel, bDof, biDimVal, dtype=None):
        self.feed.update(biInput=np.array(biInput), biLabel=np.array(biLabel), bDof=bDof, biDimVal=biDimVal)
        self.calls["upload_bic"] += 1

    def set_weights(self, w):
        self.feed["w"] = np.asarray(w, dtype=np.float64).reshape(3).copy()

    def _run(self, need_grad):
        return go.loss_and_grad(self.theta.astype(np.float32), self.feed, need_grad=need_grad, **self.kw)

    def loss(self, lossVec=False):
        self.calls["loss"] += 1
        r = self._run(False)
        out = dict(loss=np.float32(r["loss"]), BCloss=np.float32(r["BCloss"]), ICloss=np.float32(r["ICloss"]),
                   varLoss=np.float32(r["varLoss"]))
        if lossVec:
            out["lossVec"] = r["lossVec"].astype(np.float32)
        return out

    def loss_grad(self, fetch=True):
        self.calls["loss_grad"] += 1
        r = self._run(True)
        self.gbuf[:self.nparam] = r["grad"]
        self.gbuf[self.nparam:] = [r["loss"], r["BCloss"], r["ICloss"], r["varLoss"]]
        if fetch:
            return dict(loss=np.float32(r["loss"]), BCloss=np.float32(r["BCloss"]), ICloss=np.float32(r["ICloss"]),
                        varLoss=np.float32(r["varLoss"]), grad=r["grad"].astype(np.float32))

    def loss_grad_fed(self, Input, gcoef, source, N, dNt, intShape, integW, detJ, detJvec=False, dtype=None, fetch=False):
        # same contract as Engine.loss_grad_fed: upload_points + loss_grad in one call
        self.calls["loss_grad_fed"] += 1
        self.upload_points(Input, gcoef, source, N, dNt, intShape, integW, detJ, detJvec, dtype=dtype)
        r = self.loss_grad(fetch=True)
        return {k: r[k] for k in ("loss", "BCloss", "ICloss", "varLoss")} if fetch else None

    def torch_device(self):
        return "cpu"

    def synchronize(self):
        pass

    def grad_tensor(self):
        import torch
        return torch.from_numpy(self.gbuf)          # shares memory: an all-reduce lands in self.gbuf

    def optimizer_step(self, lr):
        self.calls["step"] += 1
        g = self.gbuf[:self.nparam]
        self.step += 1
        if self.optimizer.startswith("rms"):
            self.theta, self.v, self.m = go.rmsprop_step(self.theta, g, self.v, self.m, lr=lr)
        else:
            self.theta, self.m, self.v = go.adam_step(self.theta, g, self.m, self.v, self.step, lr=lr)

    def train_step(self, lr, fetch_loss=True):
        self.loss_grad(fetch=False)
        loss = self.gbuf[self.nparam]
        self.optimizer_step(lr)
        return np.float32(loss) if fetch_loss else None

    def train_batches(self, lr, tf_index):
        out = []
        for row in np.asarray(tf_index):
            self.set_batch(row)
            out.append(self.train_step(lr))
        return np.asarray(out, dtype=np.float32)

    def train_batches_begin(self, lr, tf_index):
        pend = self.__dict__.setdefault("_pending", [])
        if len(pend) >= 2:
            raise RuntimeError("two vn_train_batches_begin calls are in flight")
        pend.append(self.train_batches(lr, tf_index))

    def train_batches_end(self):
        pend = self.__dict__.setdefault("_pending", [])
        if not pend:
            raise RuntimeError("no train_batches_begin call is in flight")
        return pend.pop(0)

    def eval(self, X):
        self.calls["eval"] += 1
        X = np.asarray(X, dtype=np.float32).astype(np.float64).reshape(-1, self.inpDim)
        return go.mlp_value(self.theta.astype(np.float32).astype(np.float64), X, self.inpDim, self.layerWidth,
                            go.act_id(self.kw["activation"])).astype(np.float32)

    def residual(self, X, diff, vel, diff_dx, source):
        n = np.asarray(X).reshape(-1, self.inpDim).shape[0]
        bc = lambda a, c: np.broadcast_to(np.asarray(a, dtype=np.float64).reshape(-1, c) if np.size(a) > 1 else np.full((1, c), float(np.asarray(a).reshape(-1)[0])), (n, c))
        u, r = go.strong_residual(self.theta.astype(np.float32), X, bc(diff, 1), bc(vel, self.dim), bc(diff_dx, self.dim),
                                  bc(source, 1), self.dim, self.inpDim, self.layerWidth, self.kw["activation"],
                                  self.kw["timeDependent"])
        return u.astype(np.float32), r.astype(np.float32)
