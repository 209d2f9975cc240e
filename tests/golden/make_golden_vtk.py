"""Generates tests/golden/dem_vtk_t1_*.vtk from the REFERENCE ITSELF: examples/dem.py on the 0.1 x 0.015 x 0.04 box with its
psim.vtk_output("output/dem_cpu", frequency) line kept (variant dem_vtk_t1 of oracle/build_ref.py, frequency 30), i.e. the files
runtime/vtk.hpp writes after iterations 0 and 30.  Needs /root/reference (this container only).

    python tests/golden/make_golden_vtk.py
"""
import glob
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_worker  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(ROOT, "oracle", "_ref", "output")
for f in glob.glob(os.path.join(OUT, "dem_cpu_*.vtk")):
    os.remove(f)
ref_worker.bench_many("dem_vtk_t1", 1, 33, 1)          # runs iterations 0..34 in a fresh process, cwd = oracle/_ref
for ts in (0, 30):
    for part in ("local", "ghost"):
        src = os.path.join(OUT, f"dem_cpu_{part}_{ts}.vtk")
        dst = os.path.join(HERE, f"dem_vtk_t1_{part}_{ts}.vtk")
        shutil.copyfile(src, dst)
        print(dst, os.path.getsize(dst), "bytes")
for f in glob.glob(os.path.join(OUT, "dem_cpu_*.vtk")):
    os.remove(f)

# the STOCK example (0.8 x 0.015 x 0.2 box, 18720 spheres + 2 planes), cut to 100 iterations: what runtime/vtk.hpp writes after
# iteration 100 -> tests/golden/dem_stock_local_100.vtk.gz (tests/test_gpu_examples.py runs the example FILE on this backend)
import gzip  # noqa: E402

ref_worker.bench_many("dem_stock_t1", 1, 100, 1)        # iterations 0..101: the file of iteration 100 is written at its end
src = os.path.join(OUT, "dem_cpu_local_100.vtk")
dst = os.path.join(HERE, "dem_stock_local_100.vtk.gz")
with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as gz:
    gz.write(f.read())
print(dst, os.path.getsize(dst), "bytes")
for f in glob.glob(os.path.join(OUT, "dem_cpu_*.vtk")):
    os.remove(f)
