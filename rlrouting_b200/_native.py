"""ctypes binding of the C ABI in include/rlrouting_b200.h (built from csrc/rlr_kernels.cu).

There is NO CPU fallback: if the shared library is missing or does not load, importing the product
raises.  `build()` compiles it in-tree with nvcc for sm_100a (cross-compiles without a GPU).
"""
import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HDR = os.path.join(ROOT, "include", "rlrouting_b200.h")
STEP_HDR = os.path.join(CSRC, "rlr_step.cuh")
# translation units: the host side of the C ABI + one per policy family (each instantiates the step kernel for its policy)
UNITS = ["rlr_kernels"] + ["step_" + n for n in ("shortest", "qroute", "qroute_f32", "hybridq", "mahybridq", "cq", "cdrq", "globalroute")]
SRC = os.path.join(CSRC, "rlr_kernels.cu")
SO = os.environ.get("RLR_LIB") or os.path.join(CSRC, "librlrouting_b200.so")   # RLR_LIB: experiment builds

ABI_VERSION = 4
NUM_COUNTERS = 8
CTR = {"all_packets": 0, "end_packets": 1, "drop_packets": 2, "hops": 3, "route_time": 4, "sends": 5}
POLICY_SHORTEST, POLICY_QROUTE, POLICY_HYBRIDQ, POLICY_MAHYBRIDQ, POLICY_CQ, POLICY_CDRQ, POLICY_GLOBALROUTE = 0, 1, 2, 3, 4, 5, 6
INJECT_TRACE, INJECT_PHILOX = 0, 1
STATUS_OK, STATUS_NAN_PROB, STATUS_QUEUE_OVERFLOW, STATUS_INJECT_OVERFLOW = 0, 1, 2, 3

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC"]


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> csrc/librlrouting_b200.so for sm_100a: every translation unit that is older than its sources is
    rebuilt (all of them in parallel), then the objects are linked into the shared library."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(CSRC, "obj")
    os.makedirs(objdir, exist_ok=True)
    hdr_time = max(os.path.getmtime(HDR), os.path.getmtime(STEP_HDR))
    jobs, objs = [], []
    for u in UNITS:
        src, obj = os.path.join(CSRC, u + ".cu"), os.path.join(objdir, u + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", obj, src] + (["-Xptxas", "-v"] if verbose else [])
            jobs.append((u, subprocess.Popen(cmd)))
    failed = [u for u, p in jobs if p.wait() != 0]
    if failed:
        raise RuntimeError(f"nvcc failed for {failed}")
    if jobs or not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(o) for o in objs):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO] + objs)   # (no default-arch stub)
    return SO


class Topology(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("sum_deg", C.c_int32), ("max_deg", C.c_int32),
                ("row_ptr", C.c_void_p), ("col", C.c_void_p)]


class Config(C.Structure):
    _fields_ = [("replicas", C.c_int64), ("replica_base", C.c_int64), ("policy", C.c_int32),
                ("queue_capacity", C.c_int32), ("replicas_per_block", C.c_int32),
                ("threads_per_block", C.c_int32), ("tables_in_smem", C.c_int32), ("device", C.c_int32),
                ("event_capacity", C.c_int32), ("table_dtype", C.c_int32),
                ("page_records", C.c_int32), ("pages_per_replica", C.c_int32)]


class Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "q_table", "theta", "trace", "ring", "q_head", "q_tail", "counters", "status", "next_hop",
        "distance", "choice", "topo_row_ptr", "topo_col", "heap_pos", "gr_choice_mask", "gr_distance", "choice_mask", "gr_qs_off",
        "link_sent", "event_heap", "event_count", "event_rec", "q_dead",
        "pool", "page_next", "page_free", "page_top", "q_pages")]


class RunParams(C.Structure):
    _fields_ = [("inject_mode", C.c_int32), ("add_entropy", C.c_int32),
                ("inj_events", C.c_void_p), ("inj_cap", C.c_int32), ("is_drop", C.c_int32),
                ("enlam", C.c_void_p), ("inj_thr", C.c_void_p), ("seed", C.c_uint64),
                ("lr_q", C.c_double), ("lr_p", C.c_double), ("lr_e", C.c_double),
                ("discount", C.c_double), ("discount_trace", C.c_double), ("conf_decay", C.c_double),
                ("clock0", C.c_int32), ("shortest_random", C.c_int32), ("multiway", C.c_int32), ("phase", C.c_int32),
                ("rewards", C.c_void_p),
                ("metrics", C.c_void_p), ("route_time_out", C.c_void_p), ("hop_out", C.c_void_p),
                ("metrics2", C.c_void_p), ("droprate_out", C.c_void_p),
                ("sample", C.c_void_p), ("sample_idx", C.c_void_p), ("sample_clock", C.c_void_p),
                ("sample_size", C.c_int32), ("reserved2", C.c_int32), ("sendlog", C.c_void_p),
                ("slot", C.c_int32), ("freq", C.c_int32), ("transtime", C.c_int32), ("bandwidth", C.c_int32),
                ("general", C.c_int32), ("reserved3", C.c_int32)]


EXPORTS = ("rlr_abi_version", "rlr_last_error_string", "rlr_struct_sizes", "rlr_create", "rlr_destroy", "rlr_bind_buffers",
           "rlr_launch_info", "rlr_reset", "rlr_init_tables", "rlr_build_shortest_tables",
           "rlr_build_inject_schedule", "rlr_run_steps", "rlr_reduce_stats", "rlr_trace_replay", "rlr_trace_replay_many",
           "rlr_trace_schedule", "rlr_reduce_levels", "rlr_table_stats")

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise ImportError(
            f"{SO} is missing: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; "
            "g.build()'`). rlrouting_b200 has no CPU fallback.")
    L = C.CDLL(SO)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise ImportError(f"{SO} does not export {name}")
    p = C.c_void_p
    L.rlr_abi_version.restype = C.c_int
    L.rlr_last_error_string.restype = C.c_char_p
    L.rlr_create.argtypes = [C.POINTER(Topology), C.POINTER(Config), C.POINTER(p)]
    L.rlr_destroy.argtypes = [p]
    L.rlr_bind_buffers.argtypes = [p, C.POINTER(Buffers)]
    L.rlr_launch_info.argtypes = [p] + [C.POINTER(C.c_int32)] * 4
    L.rlr_reset.argtypes = [p, p]
    L.rlr_init_tables.argtypes = [p, C.c_double, p]
    L.rlr_build_shortest_tables.argtypes = [p, C.c_int32, p]
    L.rlr_run_steps.argtypes = [p, C.c_int64, C.POINTER(RunParams), p]
    L.rlr_build_inject_schedule.argtypes = [p, C.c_int64, C.POINTER(RunParams), p]
    L.rlr_reduce_stats.argtypes = [p, p, p]
    L.rlr_reduce_levels.argtypes = [p, p, C.c_int64, p, C.c_int32, p, p, p, p]
    L.rlr_table_stats.argtypes = [p, C.c_int32, p, p, p, C.c_int32, p, p]
    L.rlr_trace_replay.argtypes = [p, p, C.c_int32, C.c_int64, C.c_double, p, p, C.c_int64, p]
    L.rlr_trace_replay_many.argtypes = [p, p, C.c_int64, C.c_int32, C.c_int64, p, p, p, C.c_int64, p, C.c_int32]
    L.rlr_trace_schedule.argtypes = [p, p, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, p, C.c_int32, p,
                                     C.c_int32]
    if L.rlr_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI mismatch: library {L.rlr_abi_version()} != binding {ABI_VERSION}")
    sizes = (C.c_int32 * 4)()
    L.rlr_struct_sizes(sizes)
    mine = [C.sizeof(Topology), C.sizeof(Config), C.sizeof(Buffers), C.sizeof(RunParams)]
    if list(sizes) != mine:
        raise ImportError(f"struct layout mismatch: library {list(sizes)} != binding {mine}")
    _lib = L
    return L


class NativeError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise NativeError(f"rlrouting_b200 native error {rc}: {lib().rlr_last_error_string().decode()}")
