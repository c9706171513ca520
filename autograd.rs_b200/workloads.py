"""The BASELINE.json configurations as data: shapes, seeded synthetic inputs and weights, and the training
loop of examples/fit_sine.rs:64-93 written against the front end.  Pure host-side set-up (NumPy RNG only):
all arithmetic of the hot path happens in the engine.

  C1  fit_sine      X[2000,3] -> Linear(3,1) -> MSE, full batch, lr 1e-3            (examples/fit_sine.rs)
  C2  mlp4096       4 x Linear(4096,4096), ReLU between, batch 8192 per GPU         (BASELINE configs[1])
  C3  lenet         im[1024,1,28,28] -> Conv2D(1,6,5) -> ReLU -> Flatten -> Linear(3456,10) -> MSE
  C5  mlp8192       8 x Linear(8192,8192), Sigmoid between, batch 8192 per GPU      (BASELINE configs[4])
"""
import numpy as np

from . import frontend as F

CONFIGS = {
    # name: (layers, width, act, rows per GPU, lr).  The literal MSE gradient is 2/B (not 2/(B*D)) times (p - t)
    # (operators.rs:247-249), i.e. D times the analytic one: lr is scaled down so wide models stay finite.
    "mlp4096": dict(layers=4, width=4096, act="relu", rows=8192, lr=1e-5),
    "mlp8192": dict(layers=8, width=8192, act="sigmoid", rows=8192, lr=1e-6),
    "mlp_small": dict(layers=3, width=256, act="relu", rows=512, lr=1e-4),
}


def kaiming_bound(shape):
    """init.rs:19-28: gain = sqrt(2/(1+5)), fan_in = shape[1] * prod(shape[2:]), bound = sqrt(3) * gain / sqrt(fan_in)."""
    fan_in = shape[1] * int(np.prod(shape[2:])) if len(shape) > 2 else shape[1]
    return float(np.sqrt(3.0) * np.sqrt(2.0 / 6.0) / np.sqrt(fan_in))


def mlp_init(layers, width, seed=0, in_features=None, out_features=None):
    """Seeded stand-in for the reference's unseeded init (layers.rs:82-89): same distributions and bounds."""
    rng = np.random.default_rng(seed)
    ws, bs = [], []
    for i in range(layers):
        fi = in_features if (i == 0 and in_features) else width
        fo = out_features if (i == layers - 1 and out_features) else width
        wb = kaiming_bound((fi, fo))
        ws.append(rng.uniform(-wb, wb, (fi, fo)).astype(np.float32))
        bb = 1.0 / np.sqrt(fo)  # fan_in of the [in,out] weight is its dim 1 (init.rs:30-45)
        bs.append(rng.uniform(-bb, bb, (1, fo)).astype(np.float32))
    return ws, bs


def mlp_batch(rows, in_features, out_features, seed=1, rank=0):
    """X, T ~ U(-1, 1); each rank draws its own shard (seed offset) so shards differ."""
    rng = np.random.default_rng(seed + 1000 * rank)
    x = rng.uniform(-1, 1, (rows, in_features)).astype(np.float32)
    t = rng.uniform(-1, 1, (rows, out_features)).astype(np.float32)
    return x, t


def build_mlp(dev, weights, biases, act):
    """Linear (+act between layers) as nn layers holding the given weights (the parity harness injects weights)."""
    layers = []
    for i, (w, b) in enumerate(zip(weights, biases)):
        layers.append(F.nn.Linear(dev, w.shape[0], w.shape[1], w=F.Tensor.from_numpy(dev, w), bias=F.Tensor.from_numpy(dev, b)))
        if act is not None and i + 1 < len(weights):
            layers.append({"relu": F.nn.ReLU, "sigmoid": F.nn.Sigmoid}[act]())
    return F.nn.Sequential(layers)


def build_lenet(dev, filt, cbias, w, b):
    """C3: Conv2D(1->Co,k) -> ReLU -> Flatten -> Linear -> (MSE outside)."""
    k, co = filt.shape[0], filt.shape[2]
    return F.nn.Sequential([
        F.nn.Conv2D(dev, 1, co, k, 1, 0, kernels=F.Tensor.from_numpy(dev, filt), bias=F.Tensor.from_numpy(dev, cbias)),
        F.nn.ReLU(),
        F.nn.Flatten(),
        F.nn.Linear(dev, w.shape[0], w.shape[1], w=F.Tensor.from_numpy(dev, w), bias=F.Tensor.from_numpy(dev, b)),
    ])


class Trainer:
    """One training iteration exactly as the reference's examples spell it (fit_sine.rs:64-93):
    forward -> MSE -> backward -> [average gradients over ranks] -> SGD step -> zero_grad."""

    def __init__(self, model, lr, dp=None, fused_overlap=True):
        self.fused_overlap = fused_overlap   # fused exchange: per layer under backward (default) or one kernel after it
        self.model = model
        self.params = model.parameters()
        self.optim = F.nn.SGD(self.params, lr)
        self.loss_layer = F.nn.MeanSquaredError()
        self.dp = dp

    def step(self, x, t):
        pred = self.model.forward([x.clone()])
        loss = self.loss_layer.forward([pred, t.clone()])
        if self.dp is not None and self.dp.fused and self.fused_overlap:
            # exchange over NVLink peer memory, layer by layer under the backward pass: the dW GEMM's epilogue pushes the gradient
            # tiles to their owners, and as soon as a layer's gradients exist its slot sum + SGD + all-gather run on side streams.
            # Only the registered parameters exist in this model, so this replaces sync_grads() and optim.step().
            self.dp.fused_begin(self.optim.lr)
            loss.backward()
            self.dp.fused_finish()
            self.model.zero_grad()
            return loss
        loss.backward()
        if self.dp is not None and self.dp.fused:
            # the same exchange as ONE kernel after backward (cross-rank barrier, slot sum, SGD, all-gather, barrier)
            self.dp.fused_sgd_step(self.optim.lr)
        else:
            if self.dp is not None:
                self.dp.sync_grads()
            self.optim.step()
        self.model.zero_grad()
        return loss

    def reserve_pool(self, x, t, factor=2.5):
        """Run one iteration, then back the device memory pool with `factor` x its high-water mark in one block.  Two steps'
        worth of activations are alive at a time (the previous step's graph is dropped after the next one is enqueued), and
        without this the stream-ordered pool keeps growing in small increments for several steps -- each growth a driver call
        of 1-100 ms on the host inside what may be a timed region."""
        dev = x.dev
        loss = self.step(x, t)
        dev.sync()
        del loss
        used_high = dev.pool_stats()[3]
        dev.pool_reserve(int(factor * used_high) + (64 << 20))
        return used_high

    def capture(self, x, t, unroll=1, warmup=2):
        """Record `unroll` consecutive iterations on (x, t) as ONE CUDA graph.  The host walks the model once, here; replay()
        then costs one driver call per `unroll` steps -- what launch-bound configurations (fit_sine: ~11 tiny kernels per
        step) need.  The warm-up steps run normally first: workspaces and the memory pool reach their steady state outside
        the capture.  Returns the loss tensor of the last captured iteration; its buffer keeps its address across replays."""
        dev = x.dev
        for _ in range(warmup):
            self.step(x, t)
        dev.sync()
        dev.capture_begin()
        try:
            loss = None
            for _ in range(unroll):
                loss = self.step(x, t)
        finally:
            graph = dev.capture_end()
        self._graph, self._graph_loss, self._unroll = graph, loss, unroll
        return loss

    def replay(self, times=1):
        """run `times` x `unroll` training iterations from the captured graph; returns the (device-resident) loss handle"""
        self._graph.launch(times)
        return self._graph_loss
