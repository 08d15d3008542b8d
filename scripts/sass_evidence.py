"""profiles/sass_<round>.txt: counts of the Blackwell-only SASS mnemonics per kernel of the built library (run HERE, no GPU needed):
python scripts/sass_evidence.py r02"""
import collections, os, re, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "bundletrack_b200", "lib", "libbundletrack_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCMMA|UTMALDG\.\dD|UTMASTG\.\dD|LDTM[.\w]*|STTM[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|SYNCS[.\w]*|FFMA2|FADD2|HMNMX2[.\w]*|LDG\.E[.\w]*\.256[.\w]*)")
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for t in pat.findall(line):
            cnt[cur][t.split(".")[0] if t.startswith(("SYNCS", "HMNMX2")) else t] += 1
out = [f"# SASS evidence, round {R} - `cuobjdump -sass bundletrack_b200/lib/libbundletrack_b200.so` (sm_100a), Blackwell-only mnemonics per kernel",
       "# UTCHMMA = tcgen05.mma kind::f16, UTMALDG = cp.async.bulk.tensor (TMA load), LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit -> mbarrier,",
       "# UTCATOMSWS = tcgen05.alloc/dealloc, SYNCS = mbarrier ops, FFMA2/FADD2 = packed f32x2 arithmetic, HMNMX2 = packed fp16 min/max, LDG...256 = 256-bit global loads", ""]
for k, c in cnt.items():
    dem = subprocess.run(["cu++filt", k], capture_output=True, text=True).stdout.strip()
    out.append(dem[:140])
    out.append("    " + "  ".join(f"{t} x{n}" for t, n in sorted(c.items())))
open(os.path.join(ROOT, "profiles", f"sass_{R}.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
