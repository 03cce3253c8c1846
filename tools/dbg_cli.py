import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unipeak_b200.synth import HG19_NAMES, HG19_SIZES, make_synthetic
wd = tempfile.mkdtemp()
i = HG19_NAMES.index("chr21")
d = make_synthetic([HG19_SIZES[i]], 1, 1000000, directional=True, seed=2026, contig_names=["chr21"])
d.write_contig_table(os.path.join(wd, "contigs.txt")); d.write_wig(os.path.join(wd, "s0.wig"), 0, "s0")
for rep in range(2):
    t0 = time.perf_counter()
    p = subprocess.Popen([os.path.join(ROOT, "bin", "regions"), "-f", "-c", "contigs.txt", "-m", "3095693983", "-o", "out.txt", "s0.wig"], cwd=wd, stderr=subprocess.PIPE,
                         env=dict(os.environ, UPB_TRACE="1", UPB_TRACE_SLOW="1"))
    for line in p.stderr:
        print("%.3f %s" % (time.perf_counter() - t0, line.decode().rstrip()[:300]))
    p.wait()
    print("total %.3f rc %d" % (time.perf_counter() - t0, p.returncode))
