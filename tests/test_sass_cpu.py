"""Static checks on the SASS of libnlc_b200.so (no GPU needed: cuobjdump reads the cubin) that guard two measured findings
of profiles/r1_rollout_pingpong.md:

* every production tcgen05 kernel issues its MMAs under elect.sync - no generic divergence loop (BRA.U.ANY) around UTCHMMA;
* the ping-pong rollout's code stays near what the instruction cache held in the measurements (97-113 KB measured good,
  166 KB measured 14-20 % slower on most SMs)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "neurallaplacecontrol_b200", "libnlc_b200.so")
PRODUCTION = ("encode_tc2_kernel", "rollout_tc2_kernel", "rollout_pp_kernel")


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(exe) or not os.path.isfile(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, timeout=600)
    if out.returncode != 0:
        pytest.skip("cuobjdump could not read the library")
    stats = collections.defaultdict(lambda: {"instr": 0, "mma": 0, "waterfall": 0})
    bodies = collections.defaultdict(list)
    name = None
    for line in out.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name is None or not re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            continue
        st = stats[name]
        st["instr"] += 1
        body = bodies[name]
        body.append(line)
        if "UTCHMMA" in line:
            st["mma"] += 1
        m = re.search(r"BRA\.U\.ANY\s+0x([0-9a-f]+)", line)
        if m:
            # a divergence ("waterfall") loop: ptxas wraps an instruction whose uniform operand it could not prove uniform.
            # Around the one-off TMEM allocate / free (UTCATOMSWS) at the kernel's ends it is harmless; around an MMA or
            # a bulk copy it serialises the issue
            target = int(m.group(1), 16)
            loop = [b for b in body if int(re.match(r"\s+/\*([0-9a-f]+)\*/", b).group(1), 16) >= target]
            if any(op in b for b in loop for op in ("UTCHMMA", "UTCQMMA", "UBLKCP", "UTMALDG")):
                st["waterfall"] += 1
    return stats


def test_tcgen05_kernels_issue_under_elect_sync(sass):
    kernels = {n: s for n, s in sass.items() if any(p in n for p in PRODUCTION)}
    assert len(kernels) >= 20, sorted(kernels)
    for n, s in kernels.items():
        assert s["mma"] > 0, n
        assert s["waterfall"] == 0, (n, s)


def test_ping_pong_rollout_fits_the_instruction_cache(sass):
    pp = {n: s for n, s in sass.items() if "rollout_pp_kernel" in n}
    assert pp
    for n, s in pp.items():
        assert s["instr"] * 16 <= 80 * 1024, (n, s["instr"] * 16 // 1024)  # 62 KB with the group-uniform L3 units (round 2)
