"""CPU: barrier protocol of the persistent GEMM (csrc/gemm_persist.inc) under a discrete-event model of mbarrier
semantics with random latencies (tools/sim_persist_protocol.py) — no deadlock, no operand-ring or accumulator reuse
hazard, every tile accumulated and drained once."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_persistent_gemm_barrier_protocol():
    spec = importlib.util.spec_from_file_location("sim_persist_protocol", os.path.join(ROOT, "tools", "sim_persist_protocol.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(600, seed=7) == 600
