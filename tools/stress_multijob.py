"""Stress: five concurrent jobs (MLR, NMF, LDA, Lasso, GBT) on a 4-executor job server, repeated; a watchdog dumps every
thread's harmony_b200 frames and exits 3 when an iteration exceeds 60 s (hang detector for the control plane).

    python tools/stress_multijob.py [SchedulerImpl|FifoExclusiveScheduler|PackingScheduler] [iterations]
    HB_STRESS_EXECUTOR_MODE=process HB_STRESS_BACKEND=host python tools/stress_multijob.py ...   (one process per executor)
"""
import sys, logging, threading, time, traceback, os, tempfile, socket
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
logging.basicConfig(level=logging.ERROR)
from harmony_b200.jobserver.client import JobServer
from harmony_b200.dolphin.launcher import DolphinJobLauncher
from harmony_b200.mlapps import nmf, mlr, lda, lasso, gbt
from harmony_b200.common.samples import write_all
from harmony_b200.utils.flags import Configuration
deadline=[time.time()+60]
def dog():
    while True:
        time.sleep(1)
        if time.time() > deadline[0]:
            names={t.ident:t.name for t in threading.enumerate()}
            for tid, fr in sys._current_frames().items():
                st=traceback.extract_stack(fr)
                hb=[f"{g.filename.split('/')[-1]}:{g.lineno}:{g.name}" for g in st if 'harmony_b200' in g.filename]
                if hb and not any(x in hb[-1] for x in ("transport.py:4", "remote_access.py:1", "_sender_loop", "metric.py")):
                    print("THREAD", names.get(tid), " > ".join(hb[-4:]), flush=True)
            os._exit(3)
threading.Thread(target=dog, daemon=True).start()
tmp=write_all(tempfile.mkdtemp())
sched=sys.argv[1] if len(sys.argv)>1 else "SchedulerImpl"
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 6):
    s=socket.socket(); s.bind(("127.0.0.1",0)); port=s.getsockname()[1]; s.close()
    P=["-port",str(port)]+[x for x in os.environ.get("HB_STRESS_EXTRA","").split(",") if x]   # e.g. -mini_batch_grant_size,3
    subs=[(mlr,["-input",tmp+"/sample_mlr","-classes","10","-features","784","-max_num_epochs","2","-num_mini_batches","8","-num_worker_blocks","8","-step_size","0.01"]),
      (nmf,["-input",tmp+"/sample_nmf","-rank","8","-features","100","-step_size","0.01","-max_num_epochs","3","-num_mini_batches","8","-num_worker_blocks","8"]),
      (lda,["-input",tmp+"/sample_lda","-num_topics","10","-num_vocabs","17935","-max_num_epochs","2","-num_mini_batches","8","-num_worker_blocks","8"]),
      (lasso,["-input",tmp+"/sample_lasso","-features","10","-features_per_partition","5","-lambda","0.5","-step_size","1.0","-max_num_epochs","3","-num_mini_batches","4","-num_worker_blocks","4"]),
      (gbt,["-input",tmp+"/sample_gbt","-metadata_path",tmp+"/sample_gbt.meta","-features","784","-step_size","0.1","-max_num_epochs","1","-num_mini_batches","4","-num_worker_blocks","4","-max_depth_of_tree","2"])]
    deadline[0]=time.time()+60
    js=JobServer(Configuration({"num_executors":4,"port":port,"scheduler":sched,"backend":os.environ.get("HB_STRESS_BACKEND","auto"),
                                "executor_mode":os.environ.get("HB_STRESS_EXECUTOR_MODE","inproc")})).start()
    t=time.time(); names=[]
    for i,(app,a) in enumerate(subs):
        r=DolphinJobLauncher.submit_job(app.APP_PATH, app.APP, a+P, port=port); names.append(r.split("JobId: ")[-1].strip())
    oks=[js.driver.wait_job(n, timeout=600)[0] for n in names]
    print("ITER", it, sched, oks, round(time.time()-t,1), flush=True)
    js.stop(kill=True)
os._exit(0)
