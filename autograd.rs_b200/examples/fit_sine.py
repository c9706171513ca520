"""examples/fit_sine.rs (reference lines 17-102), line for line, on the B200 engine: fit y = sin(x) with the third-order polynomial
y = a + b x + c x^2 + d x^3, learned as Linear(3, 1) on the features [x, x^2, x^3] with MSELoss and full-batch SGD.

    python autograd.rs_b200/examples/fit_sine.py [--epochs 2000] [--graph]

--graph replays the training step from a CUDA graph (10 steps per launch): this configuration runs ~10 tiny kernels per step and is
bound by launch latency, not by the device.  Needs an sm_100 GPU (there is no CPU path)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import autograd_rs_b200  # noqa: F401,E402
from autograd_rs_b200 import frontend as F  # noqa: E402
from autograd_rs_b200.frontend import Tensor, nn  # noqa: E402


class MyModel(nn.NN):                                   # fit_sine.rs:39-56
    def __init__(self, dev):
        self.layer = nn.Linear(dev, 3, 1)

    def layers(self):
        return [self.layer]

    def forward(self, xs):
        return self.layer.forward(xs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2000)
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    dev = F.Device(0)
    # generate inputs and the respective sin (fit_sine.rs:21-33)
    x = np.linspace(-np.pi, np.pi, 2000).astype(np.float32)
    y = np.sin(x).reshape(-1, 1)
    xx = np.stack([x, x ** 2, x ** 3], axis=1).astype(np.float32)
    x, y = Tensor.from_numpy(dev, xx), Tensor.from_numpy(dev, y)
    y.requires_grad = False

    model = MyModel(dev)
    params = model.parameters()
    lr, n_epochs = 1e-3, a.epochs
    optim = nn.SGD(params, lr)
    loss_fn = nn.MeanSquaredError()

    def step():
        pred = model.forward([x.clone()])
        loss = loss_fn.forward([pred.clone(), y.clone()])
        loss.backward()                                 # no need to zero_grad before backward(): grads are initialised
        optim.step()
        model.zero_grad()                               # or optim.zero_grad(model)
        return pred, loss

    if a.graph:
        step()
        step()
        dev.sync()
        dev.capture_begin()
        for _ in range(10):
            pred, loss = step()
        graph = dev.capture_end()
        for epoch in range(0, n_epochs, 100):
            graph.launch(10)
            print(f"Epoch {epoch + 99} || Pred {pred.shape()} || loss {loss.data().reshape(-1)}")
        graph.close()
    else:
        for epoch in range(n_epochs):
            pred, loss = step()
            if epoch % 100 == 99:
                print(f"Epoch {epoch} || Pred {pred.shape()} || loss {loss.data().reshape(-1)}")
    w, bias = model.parameters()
    b, c, d = w.data().reshape(-1)
    print(f"Result: y = {bias.data()[0, 0]} + {b} x + {c} x^2 + {d} x^3")
    del x, y, params, optim, model, w, bias, pred, loss
    dev.close()


if __name__ == "__main__":
    main()
