"""Campaign tooling: the on-device injector emits the reference supervisor's JSON log schema."""
import json
import os
import sys

import numpy as np
import pytest

REF_SIM = "/root/reference/simulation/platform"


def test_host_plan_matches_oracle(built_lib, oracle):
    from coast_b200 import campaign as cp
    for kernel, ub, K, nc in ((oracle.K_SHA256, 64, 0, 3), (oracle.K_CRC16, 13, 0, 3), (oracle.K_AES128, 16, 0, 2), (oracle.K_MM_U32, 0, 37, 3)):
        act, rep, site, bit = cp.plan_faults(kernel, nc, ub, K, 500, seed=0x1234567890, threshold=1 << 31, unit_base=7)
        plan = oracle.make_plan(oracle.PLAN_BERNOULLI, seed=0x1234567890, threshold=1 << 31)
        for u in range(500):
            f = oracle.fault_for_unit(plan, kernel, nc, ub, K, 7 + u)
            assert (f is not None) == bool(act[u])
            if f is not None:
                assert f == (int(rep[u]), int(site[u]), int(bit[u]))
    assert cp.site_name(oracle.K_SHA256, 64, 536 + 16 + 8 * 12 + 4) == "sha256.blk1.round12.e"
    assert cp.site_name(oracle.K_AES128, 16, 16 + 16 * 9 + 15) == "aes.round9.state[15]"


@pytest.mark.skipif(not os.path.isdir(REF_SIM), reason="reference checkout absent (GPU box)")
def test_log_loads_in_the_reference_jsonparser(built_lib, tmp_path):
    """simulation/platform/jsonParser.py readJsonFile + summarizeRuns run UNCHANGED on a log written here."""
    from coast_b200 import campaign as cp
    recs = []
    for u in range(10):
        if u < 6:
            res = cp.run_result(0, 0, 1e-6)                 # success
        elif u < 8:
            res = cp.run_result(0, 3, 1e-6)                 # fault (TMR corrected)
        elif u < 9:
            res = cp.run_result(1, 0, 1e-6)                 # error (SDC)
        else:
            res = cp.abort_result("FAULT_DETECTED_DWC")     # DWC detection -> abort
        recs.append(cp.injection_record(u, "registers", f"replica0:site{u}", 0, 1 << u, f"site{u}", res, cycles=u))
    path = str(tmp_path / "campaign.json")
    cp.write_log(path, cp.R.lib_path(), recs)
    import types
    for missing in ("matplotlib", "matplotlib.pyplot", "elftools", "elftools.elf", "elftools.elf.elffile",
                    "elftools.elf.sections", "elftools.elf.descriptions", "elftools.elf.constants"):
        try:                                                # third-party plotting / ELF packages the parser imports but the
            __import__(missing)                             # summary path never calls; absent in this image
        except ImportError:
            sys.modules[missing] = types.ModuleType(missing)
    for mod, names in (("elftools.elf.elffile", ["ELFFile"]), ("elftools.elf.sections", ["SymbolTableSection"]),
                       ("elftools.elf.descriptions", ["describe_sh_flags"]), ("elftools.elf.constants", ["SH_FLAGS"])):
        for nm in names:
            if not hasattr(sys.modules[mod], nm):
                setattr(sys.modules[mod], nm, object)
    sys.path.insert(0, REF_SIM)
    try:
        import jsonParser                                   # the reference's own parser
        runs, exe = jsonParser.readJsonFile(path)
        summary = jsonParser.summarizeRuns(runs, "b200")
    finally:
        sys.path.remove(REF_SIM)
    assert exe == cp.R.lib_path() and len(runs) == 10
    assert (summary.success, summary.faults, summary.errors, summary.timeouts, summary.aborts) == (6, 2, 1, 1, 1)


@pytest.mark.gpu
def test_campaigns_reproduce_the_table_shape(rt, oracle, tmp_path):
    """5 000 single-bit injections per cell, as in docs/source/results/msp430.rst -- one launch each."""
    from coast_b200 import campaign as cp
    n = 5000
    unmit, _ = cp.run_campaign(rt, "crc16", "", n, seed=3)
    dwc, _ = cp.run_campaign(rt, "crc16", "-DWC", n, seed=3)
    tmr, recs = cp.run_campaign(rt, "crc16", "-TMR -countErrors", n, seed=3, log_path=str(tmp_path / "crc_tmr.json"))
    # every flip of a live crc/data value propagates (the CRC update is a bijection of its state)
    assert unmit.errors == n and unmit.success == 0
    assert dwc.detected == n and dwc.errors == 0
    assert tmr.errors == 0 and tmr.faults == n                 # all corrected
    with open(tmp_path / "crc_tmr.json") as f:
        assert f.readline().strip().endswith("libcoast_rt.so")
        data = json.load(f)
    assert len(data) == n and set(data[0]) >= {"timestamp", "number", "section", "oldValue", "newValue", "address", "sleepTime",
                                               "cycles", "PC", "name", "result", "cacheInfo"}
    # SHA-256 and AES: TMR never lets a single flip through; unmitigated SDC rate == fraction of flips that matter (all, here)
    q3, _ = cp.run_campaign(rt, "qsort", "-TMR -countErrors", 2000, seed=5)
    q2, _ = cp.run_campaign(rt, "qsort", "-DWC", 2000, seed=5)
    q1, _ = cp.run_campaign(rt, "qsort", "", 2000, seed=5)
    # quicksort is the workload where most flips are MASKED (a compare operand flip rarely changes the branch): like the
    # published table, unmitigated still sorts correctly most of the time; TMR never fails; DWC flags a superset of the SDCs
    assert q3.errors == 0 and q1.errors < 1000 and q1.success > 1000 and q2.errors <= 2 and q2.detected >= q1.errors * 0.5
    for wl in ("sha256", "aes", "mm", "chsha"):
        t, _ = cp.run_campaign(rt, wl, "-TMR -countErrors", 2000, seed=5)
        assert t.errors == 0
        d, _ = cp.run_campaign(rt, wl, "-DWC", 2000, seed=5)
        u, _ = cp.run_campaign(rt, wl, "", 2000, seed=5)
        assert d.errors == 0 and d.detected + d.success == 2000
        # detected under DWC == SDC when unprotected: the same flips, the same propagation (replica index differs, value path does not)
        assert abs(d.detected - u.errors) <= 2000 * 0.02
