"""Build tuning variants of libcoflux.so HERE (nvcc cross-compiles) and time them on the GPU box.

    python tools/ab_variants.py build            # → climaocean.jl_b200/lib/variants/<name>.so (travel with gpurun)
    python tools/ab_variants.py run [bits] [cfg…] # on the GPU box: tools/quick_bench.py once per variant (COFLUX_LIB)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "climaocean.jl_b200", "lib", "variants")
VARIANTS = {          # name -> -D definitions; edit for the experiment at hand (results: profiles/README.md)
    "base": [],
    "psi_sm_16_binades": ["COFLUX_PSI_SM_KLO=-8", "COFLUX_PSI_SM_KHI=8"],     # 3.015 vs 3.045 ms (`:default`), but `:corrected` drops to 1 CTA/SM
}


def build():
    import importlib.util
    spec = importlib.util.spec_from_file_location("coflux_build", os.path.join(ROOT, "climaocean.jl_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    os.makedirs(VDIR, exist_ok=True)
    for f in os.listdir(VDIR):
        os.remove(os.path.join(VDIR, f))
    procs = []
    for name, defs in VARIANTS.items():
        out = os.path.join(VDIR, name + ".so")
        cmd = [b.NVCC] + b.NVCC_FLAGS + [f"-D{d}" for d in defs] + ["-Xptxas", "-v", "-o", out, b.SRC]
        procs.append((name, subprocess.Popen(cmd, stderr=open(os.path.join(VDIR, name + '.log'), 'w'))))
    for name, p in procs:
        print(name, "rc", p.wait())


def run(args):
    for name in sorted(os.listdir(VDIR)):
        if not name.endswith(".so"):
            continue
        env = dict(os.environ, COFLUX_LIB=os.path.join(VDIR, name))
        print(f"--- {name}", flush=True)
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", os.environ.get("AB_SCRIPT", "quick_bench.py"))] + args, env=env)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(sys.argv[2:])
