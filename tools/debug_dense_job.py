"""Debug helper: MLR job on the device backend, sparse (pull) vs -dense_input (tcgen05), losses of a few runs."""
import os, socket, sys, tempfile, logging
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
logging.disable(logging.CRITICAL)
from helpers import write_mlr_sample
from harmony_b200.dolphin.launcher import DolphinJobLauncher
from harmony_b200.jobserver.client import JobServer
from harmony_b200.mlapps import mlr
from harmony_b200.utils.flags import Configuration
tmp = tempfile.mkdtemp(prefix="hbgpu4")
data = write_mlr_sample(os.path.join(tmp, "sample_mlr"), n=256, features=64, classes=4, nnz=32)
s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
js = JobServer(Configuration({"num_executors": 2, "port": port, "backend": "device"})).start()
try:
    i = 0
    for dense in ("true", "false", "true", "false", "true"):
        args = ["-input", data, "-max_num_epochs", "10", "-num_mini_batches", "4", "-num_worker_blocks", "4",
                "-num_server_blocks", "8", "-step_size", "0.05", "-classes", "4", "-features", "64",
                "-dense_input", dense, "-port", str(port)]
        DolphinJobLauncher.submit_job(mlr.APP_PATH, mlr.APP, args, port=port)
        ok, result = js.driver.wait_job(f"MLR-{i}", timeout=60); i += 1
        print("dense", dense, ok, [round(r["last_batch_loss"] / 64.0, 4) for r in result] if ok else result, flush=True)
finally:
    js.stop(kill=True)
