"""Host-side check of the STFT kernel's shared-memory exchange layout (styler_b200/csrc/stft.cu, `slot()`): the XOR swizzle must
be a permutation of the 512 complex slots and every 8-byte access pattern of the three radix-8 passes must touch 16 distinct
banks per half-warp (a shared-memory wavefront serves one half-warp of 8-byte accesses: 16 banks of 8 bytes).  The formulas
below restate the index arithmetic of the kernel; the additive skew of the first version (i + (i >> 4)) is kept as the
counter-example whose pass-2 stores were two-way conflicted (profiles/ncu_stft_r2l_packed.md: 28 % of the wavefronts)."""
import re
from pathlib import Path


def slot(i):
    t = (i >> 4) & 7
    return i ^ (t | ((t & 4) << 1))


def skew(i):
    return i + (i >> 4)


def wavefronts(idx, fn):
    total = 0
    for h in range(2):
        banks = {}
        for lane in range(16 * h, 16 * h + 16):
            s = fn(idx[lane])
            banks.setdefault(s % 16, set()).add(s)
        total += max(len(v) for v in banks.values())
    return total


def patterns():
    pats = {"pass1_store": [], "strided_load": [], "pass2_store": [], "pass3_store": []}
    for u in range(2):
        for r in range(8):
            pats["pass1_store"].append([(lane + 32 * u) * 8 + r for lane in range(32)])
            pats["strided_load"].append([(lane + 32 * u) + 64 * r for lane in range(32)])
            for name, ns in (("pass2_store", 8), ("pass3_store", 64)):
                idx = []
                for lane in range(32):
                    j = lane + 32 * u
                    k = j & (ns - 1)
                    idx.append((((j - k) << 3) + k) + r * ns)
                pats[name].append(idx)
    return pats


def test_swizzle_is_a_permutation():
    assert sorted(slot(i) for i in range(512)) == list(range(512))


def test_every_pass_is_conflict_free():
    for name, accesses in patterns().items():
        for idx in accesses:
            assert sorted(idx) == sorted(set(idx)) and 0 <= min(idx) and max(idx) < 512, name
            assert wavefronts(idx, slot) == 2, name          # one wavefront per half-warp


def test_additive_skew_was_conflicted_in_pass2():
    w = sum(wavefronts(idx, skew) for idx in patterns()["pass2_store"])
    assert w == 64                                           # 16 stores x 4 wavefronts instead of 2


def test_kernel_source_uses_this_swizzle():
    src = (Path(__file__).resolve().parents[1] / "styler_b200" / "csrc" / "stft.cu").read_text()
    m = re.search(r"auto slot = \[\]\(int i\) \{(.*?)\};", src)
    assert m is not None
    body = re.sub(r"\s+", " ", m.group(1)).strip()
    assert body == "const int t = (i >> 4) & 7; return i ^ (t | ((t & 4) << 1));", body
