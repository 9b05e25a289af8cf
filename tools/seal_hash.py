import sys, hashlib
sys.path.insert(0, '/root/repo')
import hfb200_loader
pkg = hfb200_loader.load()
lib = pkg.load_library(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1] else pkg.load_library()
with pkg.Context(0, 20, (16, 192, 48), lib=lib) as c:
    c.witgen_synth(20, 77, 1)
    print(hashlib.sha256(c.prove_resident(5).tobytes()).hexdigest())
