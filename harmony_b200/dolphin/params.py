"""Dolphin / job-server / app flags with the reference's short names (CLI compatibility surface).

Reference: jobserver/.../dolphin/DolphinParameters.java:26-263, common/.../param/Parameters.java:25-119,
jobserver/.../jobserver/Parameters.java:25-94, optimizer/parameters/*.java, dashboard/parameters/*.java,
per-app *Parameters.java.  Flags that are meaningless on a GPU box
(-heap_slack, -driver_memory, ...) are accepted and ignored.
"""
from __future__ import annotations

from ..utils.flags import Param, ParamRegistry

P = Param
# ---- common client / driver (CM/param/Parameters.java) ----
COMMON = [
    P("driver_memory", int, 256, "ignored", "client"), P("eval_size", int, 128, "ignored", "client"),
    P("num_eval_cores", int, 1, "ignored", "client"), P("input", str, None, "input data path", "worker"),
    P("local", bool, True, "run on this box", "client"), P("split", int, 1, "number of input splits", "client"),
    P("timeout", int, 100000, "client timeout (ms)", "client"),
    P("max_num_eval_local", int, 2, "max local executors", "client"),
    P("heap_slack", float, 0.0, "ignored (JVM heap slack)", "client"),
    P("opt_benefit_threshold", float, 0.0, "min relative benefit to apply a plan", "optimizer"),
    P("hyper_thread_enabled", bool, False, "halve trainer threads", "worker"),
    P("default_bandwidth", float, 1.074e9, "bits/s assumed per host (NVLink when on GPUs)", "optimizer"),
    P("default_core_num", int, 1, "cores assumed per host", "optimizer"),
    P("host_to_bandwidth_file_path", str, "", "host -> bandwidth file", "optimizer"),
    P("host_to_core_file_path", str, "", "host -> #cores file", "optimizer"),
]
# ---- job server (JS/jobserver/Parameters.java) ----
JOBSERVER = [
    P("port", int, 7008, "job-server command port", "client"),
    P("chkp_commit_path", str, "/tmp/chkp/commit", "checkpoint commit dir", "driver"),
    P("chkp_temp_path", str, "/tmp/chkp", "checkpoint temp dir", "driver"),
    P("num_executors", int, 1, "executors in the resource pool", "driver"),
    P("executor_mem_size", int, 128, "ignored", "driver"), P("executor_num_cores", int, 1, "ignored", "driver"),
    P("executor_num_tasklets", int, 4, "tasklet threads per executor", "driver"),
    P("handler_queue_size", int, 1024, "", "driver"), P("sender_queue_size", int, 1024, "", "driver"),
    P("handler_num_threads", int, 2, "", "driver"), P("sender_num_threads", int, 2, "", "driver"),
    P("scheduler", str, "edu.snu.cay.jobserver.driver.SchedulerImpl", "JobScheduler implementation", "driver"),
    P("kill", bool, False, "stop_jobserver: kill running jobs", "client"),
    P("backend", str, "auto", "auto | host | device (table/compute backend)", "driver"),
    P("executor_mode", str, "inproc", "inproc | process (one OS process per executor/GPU)", "driver"),
]
# ---- Dolphin (DolphinParameters.java) ----
DOLPHIN = [
    P("max_num_epochs", int, 1, "", "master"), P("num_mini_batches", int, 10, "mini-batches per epoch (global)", "master"),
    P("clock_slack", int, 0, "SSP staleness bound", "master"),
    P("server_metric_flush_period_ms", int, 0, "", "master"),
    P("offline_model_eval", bool, False, "checkpoint the model every epoch, evaluate after training", "master"),
    P("model_eval", bool, False, "evaluate a checkpointed model instead of training", "master"),
    P("num_server_blocks", int, 1024, "", "server"), P("load_model", bool, False, "", "server"),
    P("model_path", str, "", "", "server"), P("local_model_path", str, "", "", "server"),
    P("num_trainer_threads", int, 0, "0 = auto", "worker"), P("num_worker_blocks", int, 1024, "", "worker"),
    P("test_data_path", str, "", "", "worker"), P("input_chkp_path", str, "", "", "worker"),
    P("model_cache_enabled", bool, False, "", "worker"),
    P("job_priority", str, "normal", "normal | high: CUDA stream priority of the job's tasklets on shared GPUs", "worker"),
    P("mini_batch_grant_size", int, 0, "mini-batches leased per SSP sync (1 = one driver round trip per mini-batch, "
                                       "as the reference; k > 1 saves k - 1 of them and widens the staleness bound by "
                                       "k - 1; leases taper to 1 as the budget runs out; ignored while the job's task-unit "
                                       "barriers are on; 0 = auto: 4 for the fused device trainers -- their kernels queue "
                                       "on the stream and the host only syncs at the end of a lease -- else 1)", "worker"),
    P("gpu_sm_share", float, 0.0, "fraction of a GPU's SMs this job's kernels are launched for (device backend); "
                                  "0 = auto: 1 / (jobs with an active worker on that GPU) -- the GPU twin of the "
                                  "resource-typed task-unit slots", "worker"),
    P("task_unit_scheduling", str, "auto", "on | off | auto: per-phase global barriers that interleave the phases of "
                                           "co-located jobs (one driver round trip per phase); auto = on unless the "
                                           "job server's scheduler runs jobs exclusively", "worker"),
    P("batch_metric_flush_interval_ms", int, 50, "batch metrics are flushed to the driver at most this often "
                                                 "(0 = after every mini-batch, as the reference)", "worker"),
    P("num_dolphin_workers", int, 1, "", "launcher"), P("num_servers", int, 1, "0 = co-located", "launcher"),
    P("dolphin_worker_mem_size", int, 128, "ignored", "launcher"), P("server_mem_size", int, 128, "ignored", "launcher"),
    P("num_dolphin_worker_cores", int, 1, "ignored", "launcher"), P("num_server_cores", int, 1, "ignored", "launcher"),
    P("num_server_handler_threads", int, 2, "", "launcher"), P("num_server_sender_threads", int, 2, "", "launcher"),
    P("num_worker_handler_threads", int, 2, "", "launcher"), P("num_worker_sender_threads", int, 2, "", "launcher"),
    P("server_handler_queue_size", int, 500000, "", "launcher"), P("server_sender_queue_size", int, 500000, "", "launcher"),
    P("worker_handler_queue_size", int, 500000, "", "launcher"), P("worker_sender_queue_size", int, 500000, "", "launcher"),
    P("features", int, 0, "", "user"), P("lambda", float, 0.0, "L2 regularisation", "user"),
    P("decay_rate", float, 0.9, "", "user"), P("decay_period", int, 5, "", "user"),
    P("step_size", float, 0.0, "", "user"), P("model_gaussian", float, 0.001, "", "user"),
    P("features_per_partition", int, 0, "", "user"),
]
OPTIMIZER = [
    P("optimizer", str, "edu.snu.cay.dolphin.optimizer.impl.EmptyPlanOptimizer", "Optimizer implementation", "optimizer"),
    P("et_plan_executor", str, "edu.snu.cay.services.et.plan.impl.PlanExecutorImpl", "", "optimizer"),
    P("optimization_interval_ms", int, 10000, "", "optimizer"), P("delay_after_optimization_ms", int, 10000, "", "optimizer"),
    P("num_initial_batch_metrics_to_skip", int, 3, "", "optimizer"),
    P("min_num_required_batch_metrics", int, 3, "", "optimizer"),
    P("moving_avg_window_size", int, 0, "", "optimizer"), P("metric_weight_factor", float, 0.0, "", "optimizer"),
    P("num_extra_resources", int, 0, "", "optimizer"), P("extra_resources_period_sec", int, 0, "", "optimizer"),
    P("dashboard_enabled", bool, False, "", "dashboard"), P("dashboard_port", int, 9005, "", "dashboard"),
]
MLR = [P("classes", int, 0, "number of classes", "user"),
       P("dense_input", bool, False, "dense instances: tcgen05 GEMM path on the device backend", "user"),
       P("mlr_mode", str, "pull", "pull: model rows move (PS pull/push); ship: activations move (owner-side dot, "
                                  "BSP lockstep, device backend, num_worker_blocks % workers == 0)", "user")]
NMF = [P("rank", int, 0, "", "user"), P("max_val", float, 1e6, "", "user"), P("init_max", float, 1.0, "", "user"),
       P("init_min", float, 0.0, "", "user"), P("print_mat", bool, False, "", "user")]
LDA = [P("alpha", float, 0.001, "", "user"), P("beta", float, 0.01, "", "user"), P("num_topics", int, 100, "", "user"),
       P("num_vocabs", int, 0, "", "user")]
GBT = [P("gamma", float, 0.0, "", "user"), P("max_depth_of_tree", int, 7, "", "user"),
       P("leaf_min_size", int, 4, "", "user"), P("metadata_path", str, "", "", "user"), P("num_keys", int, 0, "", "user")]
EXAMPLES = [P("delta", int, 1, "", "user"), P("num_keys", int, 0, "", "user"), P("compute_time_ms", int, 300, "", "user"),
            P("num_training_data", int, 0, "", "user"), P("num_test_data", int, 0, "", "user"),
            P("vector_size", int, 1, "", "user"), P("start_key", int, 0, "", "user")]
PREGEL = [P("num_pregel_workers", int, 1, "", "launcher"), P("pregel_worker_mem_size", int, 128, "ignored", "launcher"),
          P("pregel_worker_num_cores", int, 1, "ignored", "launcher"), P("num_worker_threads", int, 0, "", "worker"),
          P("source", int, 0, "SSSP source vertex", "user")]


# ---- ET examples / common services (services/et/.../examples/*, common/datastorer, common/metric, utils/trace) ----
ET_EXAMPLES = [
    P("num_workers", int, 2, "addinteger: updater executors", "user"),
    P("num_updates", int, 128, "addinteger: total updates by each worker", "user"),
    P("delta_value", int, 1, "addinteger: delta added by each update", "user"),
    P("coefficient", int, 1, "addinteger: multiplier the update function applies to the delta", "user"),
    P("metric_flush_period_ms", int, 0, "addinteger: automatic metric flush period (0 = off)", "user"),
    P("table_input", str, "", "simple: optional input file bulk-loaded into the table", "user"),
    P("kv_data_path", str, "", "load: `key value` lines for the exist-key loader", "user"),
    P("none_key_data_path", str, "", "load: value lines for the none-key loader", "user"),
    P("metric_automatic_flush_period_ms", int, 200, "metric: automatic flush period", "user"),
    P("metric_manual_flush_period_ms", int, 100, "metric: manual flush period inside the tasklet", "user"),
    P("custom_metric_record_period_ms", int, 20, "metric: period of the custom timestamp metric", "user"),
    P("task_duration_ms", int, 400, "metric: tasklet run time", "user"),
    P("inputs", str, "", "linecount: comma-separated input paths (dataloader example)", "user"),
]
# ---- shared services: trace (utils/trace/parameters), datastorer base dir, memory tracker ----
SERVICES = [
    P("base_dir", str, "/tmp/harmony_b200", "datastorer base directory (also receives trace files)", "driver"),
    P("memory_measure_interval", int, 1000, "memory metric tracker sampling period (ms)", "worker"),
    P("htrace_receiver_type", str, "", "'' = tracing off | STDOUT | CHROME (chrome://tracing JSON under base_dir) | "
                                       "ZIPKIN (JSON spans POSTed to host:port)", "client"),
    P("htrace_receiver_host", str, "localhost", "span receiver host (ZIPKIN)", "client"),
    P("htrace_receiver_port", int, 9410, "span receiver port (ZIPKIN)", "client"),
]


def registry(*groups) -> ParamRegistry:
    reg = ParamRegistry()
    for g in groups:
        reg.register_all(g)
    return reg


def dolphin_registry(*app_groups) -> ParamRegistry:
    return registry(COMMON, JOBSERVER, DOLPHIN, OPTIMIZER, SERVICES, *app_groups)
