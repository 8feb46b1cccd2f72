"""Debug helper: the Lasso job on the device backend, started from a thread whose current device is `dev`."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from harmony_b200.dolphin.launcher import ETDolphinLauncher
    from harmony_b200.mlapps import lasso
    tmp = tempfile.mkdtemp(prefix="hbk12")
    rng = np.random.default_rng(1)
    F = 12
    w = np.zeros(F); w[[1, 5, 9]] = [2.0, -3.0, 1.5]
    path = os.path.join(tmp, "sample_lasso")
    with open(path, "w") as f:
        for _ in range(256):          # 4 even blocks of 64 (> F): a tiny first batch could zero a support
            x = rng.standard_normal(F)     # weight, which the reference's skip-zero rule then never revisits
            f.write(f"{x @ w + 0.01 * rng.standard_normal():.5f} " + " ".join(f"{i}:{v:.5f}" for i, v in enumerate(x)) + "\n")
    from harmony_b200.mlapps import lasso_device
    orig = lasso_device.LassoDeviceTrainer.push_update
    hist = []

    def push_update(self):
        orig(self)
        hist.append((round(float(self.stats[0]), 4), tuple(self._Xt.shape), round(float(self._Xt.abs().sum()), 2),
                     round(float(self._y.abs().sum()), 2), round(float((self._Xt.double() @ self._y.double()).abs().sum()), 2), [round(float(x), 3) for x in self.w_new.tolist()],
                     [round(float(x), 3) for x in self.w_old.tolist()]))
    lasso_device.LassoDeviceTrainer.push_update = push_update
    for dev in [int(a) for a in sys.argv[1:]] or [0]:
        torch.cuda.set_device(dev)
        for rep in range(3):
            res = ETDolphinLauncher.launch(lasso.APP, ["-input", path, "-max_num_epochs", "10", "-num_mini_batches", "4",
                                                       "-num_worker_blocks", "4", "-num_server_blocks", "2", "-features", str(F),
                                                       "-features_per_partition", "6", "-step_size", "1.0", "-lambda", "0.5",
                                                       "-model_gaussian", "0.1", "-num_dolphin_workers", "1",
                                                       "-num_servers", "0", "-decay_rate", "1.0", "-backend", "device"])
            print('cpus', os.cpu_count(), 'gpus', torch.cuda.device_count())
            for h in hist[:5] + hist[-2:]:
                print("   step", h, flush=True)
            hist.clear()
            print("main-thread device", dev, "rep", rep, "->", res.ok, res.results, "current now", torch.cuda.current_device(),
                  flush=True)


if __name__ == "__main__":
    main()
