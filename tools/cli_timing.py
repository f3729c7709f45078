"""Wall-clock of the C++ host program on a named workload: FASTQ -> <prefix>.reads + <prefix>.graph3.
python tools/cli_timing.py [cfg2|cfg1|...]   (writes the FASTQ under /tmp first, not timed)"""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sage2_b200 import api, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reads, k = synth.config(name)
d = tempfile.mkdtemp(prefix="cli_")
fq = os.path.join(d, "in.fastq")
t = time.perf_counter(); synth.write_fastq(fq, reads); print(f"wrote {os.path.getsize(fq) / 1e6:.0f} MB FASTQ in {time.perf_counter() - t:.1f} s", flush=True)
binp = os.path.join(os.path.dirname(api.LIB_PATH), "sage2gpu")
for rep in range(2):
    t = time.perf_counter()
    subprocess.run([binp, "-f", fq, "-k", str(k), "-o", os.path.join(d, "out"), "-p", "g", "-M", "3"], check=True)
    print(f"run {rep}: sage2gpu -M 3 wall {time.perf_counter() - t:.2f} s", flush=True)
log = open(os.path.join(d, "out", "g.log")).read()
print("\n".join(l for l in log.splitlines() if " sec." in l or "Device ms" in l))
for ext in (".reads", ".graph3"):
    print(ext, os.path.getsize(os.path.join(d, "out", "g" + ext)) / 1e6, "MB")
