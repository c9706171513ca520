"""examples/example.rs (reference lines 11-76), line for line, on the B200 engine: a 784 -> 128 -> Sigmoid -> 10 model defined by
implementing `NN`, trained on two batches of 8 with MSELoss and the optimiser step written out by hand on the parameter tensors
(`*w -= &(grad * lr)`: in place on the shared storage, so every clone of the parameter sees the update).  Afterwards the weights
are saved and loaded back (model serialisation is on the reference's TODO list, README.md:71).

    python autograd.rs_b200/examples/example.py          (needs an sm_100 GPU: there is no CPU path)"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import autograd_rs_b200  # noqa: F401,E402
from autograd_rs_b200 import frontend as F  # noqa: E402
from autograd_rs_b200.frontend import Tensor, nn  # noqa: E402


class MyModel(nn.NN):                                   # example.rs:12-32
    def __init__(self, dev):
        self.layer1 = nn.Linear(dev, 784, 128)
        self.sig = nn.Sigmoid()
        self.layer2 = nn.Linear(dev, 128, 10)

    def layers(self):
        return [self.layer1, self.sig, self.layer2]

    def forward(self, xs):
        x = self.layer1.forward(xs)
        x = self.sig.forward([x])
        return self.layer2.forward([x])


def main():
    dev = F.Device(0)
    model = MyModel(dev)
    # data (example.rs:36-37: zeros; a seeded draw here so that the loss moves)
    rng = np.random.default_rng(0)
    x_train = rng.uniform(0, 1, (16, 784)).astype(np.float32)
    y_train = rng.uniform(0, 1, (16, 10)).astype(np.float32)
    lr, n_epochs, batch_size = 0.1, 10, 8
    mse = nn.MeanSquaredError()
    for epoch in range(n_epochs):
        for i in range(0, 16, batch_size):
            x = Tensor.from_numpy(dev, x_train[i:i + batch_size])
            y = Tensor.from_numpy(dev, y_train[i:i + batch_size])
            y.requires_grad = False
            pred = model.forward([x])
            loss = mse.forward([pred.clone(), y.clone()])
            print(f"Epoch {epoch} || Pred {pred.shape()} || loss {loss.data().reshape(-1)}")
            assert not y.has_grad()                     # no grad is computed for y
            loss.backward()
            # ** optimizer step **  (example.rs:62-71)
            for p in model.parameters():
                w_grad_scaled = p.grad() * lr           # &grad * lr -> new tensor
                p -= w_grad_scaled                      # in place on the shared storage: all clones have the updated weights
            model.zero_grad()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "model.agrs")
        model.save(path)
        other = MyModel(dev)
        other.load(path)
        same = all(np.array_equal(p.data(), q.data()) for p, q in zip(model.parameters(), other.parameters()))
        print(f"saved {os.path.getsize(path)} bytes, loaded into a fresh model: identical = {same}")
        assert same
    del model, other, x, y, pred, loss, p, w_grad_scaled
    dev.close()


if __name__ == "__main__":
    main()
