"""Dumps the BVH + triangles of a bench scene, builds traversal_whatif.cpp and runs it (CPU only).
  python tools/whatif/run.py bunny      # BASELINE config 2 geometry (82 k triangles)
  python tools/whatif/run.py soup       # config 3 (1 M triangles)"""
import os, subprocess, sys, tempfile
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
from fspt_b200 import scenes

which = sys.argv[1] if len(sys.argv) > 1 else "bunny"
sa, cam = (scenes.bunny_class(subdiv=6, atlas_res=64, env_size=(128, 64)) if which == "bunny"
           else scenes.sphere_soup(env_size=(128, 64)))
tmp = tempfile.mkdtemp()
sa.bvh.astype(np.float32).tofile(os.path.join(tmp, which + "_bvh.bin"))
sa.tris.astype(np.float32).tofile(os.path.join(tmp, which + "_tris.bin"))
exe = os.path.join(tmp, "whatif")
subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-w", "-o", exe, os.path.join(HERE, "traversal_whatif.cpp")])
subprocess.check_call([exe, os.path.join(tmp, which)] + [str(x) for x in cam["eye"] + cam["dir"]])
