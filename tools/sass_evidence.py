"""profiles/<round>_sass_evidence.txt: per kernel, the tensor-core / TMA / tensor-memory / mbarrier / dependent-launch
instructions found in the sm_100a cubin of libc4a0_engine.so (cuobjdump -sass)."""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "c4a0_b200/libc4a0_engine.so"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_sass_evidence.txt"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)
out = ["# SASS evidence (cuobjdump -sass c4a0_b200/libc4a0_engine.so, sm_100a cubin; built by c4a0_b200/build.py)",
       "# Per kernel: counts of the tensor-core (UTCHMMA = tcgen05.mma), TMA (UTMALDG = cp.async.bulk.tensor), tensor-memory",
       "# (LDTM = tcgen05.ld, UTCATOMSWS = tcgen05.alloc/dealloc, UTCBAR = tcgen05.commit), mbarrier (SYNCS) and",
       "# programmatic-dependent-launch (ACQBULK / PREEXIT / griddepcontrol) instructions, then the matching lines in program order.", ""]
pat = re.compile(r"\b(UTCHMMA[\w.]*|UTCMMA[\w.]*|UTMALDG[\w.]*|UTMACCTL[\w.]*|LDTM[\w.]*|UTCATOMSWS[\w.]*|UTCBAR[\w.]*|SYNCS[\w.]*|ELECT[\w.]*|"
                 r"UBLKCP[\w.]*|FENCE[\w.]*|REDG?[\w.]*|ATOMG?[\w.]*|CCTL[\w.]*|ACQBULK[\w.]*|PREEXIT[\w.]*|DEPBAR[\w.]*)")
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if not any(k in name for k in ("k_net2", "5k_netE", "6k_step")):
        continue
    lines = f.split("\n")
    cnt, hits = {}, []
    for ln in lines:
        m = pat.search(ln)
        if m and "/*" in ln:
            op = m.group(1)
            cnt[op] = cnt.get(op, 0) + 1
            if op.startswith(("UTC", "UTMA", "LDTM", "ACQBULK", "PREEXIT")):
                hits.append(re.sub(r"\s+", " ", ln.strip())[:150])
    n_instr = sum(1 for ln in lines if re.search(r"/\*[0-9a-f]{4}\*/", ln))
    out.append(f"## {name}")
    out.append(f"instructions: {n_instr}")
    out.append("counts: " + ", ".join(f"{k} x{v}" for k, v in sorted(cnt.items())))
    out += hits
    out.append("")
open(dst, "w").write("\n".join(out))
print("\n".join(l for l in out if l.startswith(("##", "counts"))))
