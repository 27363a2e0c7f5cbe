"""Stage the reference's own Python sources for the CPU arm of bench.py (TEST / BASELINE INFRASTRUCTURE).

    python oracle/make_ref.py            # /root/reference -> baseline/_ref/

The reference is pure Python (no setup.py: `pip install --target baseline/_ref /root/reference` has nothing to
build), so "installing" it is copying the modules its hot path imports.  `baseline/_ref/` is git-ignored (the
history stays free of reference sources) but NOT gpurun-ignored, so the staged copy travels to the GPU box, where
/root/reference does not exist, and `bench.py --impl reference` / the `cpu_baseline` leg time the REAL reference
(`cpu_baseline.kind = "reference"`); without the staged copy they fall back to the port (oracle/ref_port.py,
`kind = "port"`).  __graft_entry__.build() runs this whenever /root/reference is present.
"""
from __future__ import annotations

import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("DIC_REFERENCE", "/root/reference")
DEST = os.path.join(REPO, "baseline", "_ref")
# the modules the timed path imports: operators, the Nets around them, the p2 sweep, their helpers
FILES = ["interpolation_layer.py", "rbf.py", "dec.py", "utils.py", "info.py", "pretrain_interp.py",
         "clustering_interp.py", "p2_clustering_optK.py", "internal_eval.py"]


def stage(verbose=True):
    if not os.path.isdir(REFERENCE):
        return None
    os.makedirs(DEST, exist_ok=True)
    for f in FILES:
        src = os.path.join(REFERENCE, f)
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(DEST, f))
    if verbose:
        print(f"staged {len(os.listdir(DEST))} reference modules in {DEST}")
    return DEST


def import_reference(p2=False):
    """Import the staged reference modules (stubbing the optional imports the container lacks).
    Returns a namespace or None when nothing is staged.  p2=True also imports p2_clustering_optK (class KM)."""
    import importlib.machinery
    import types
    if not os.path.exists(os.path.join(DEST, "interpolation_layer.py")):
        return None

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)      # torch._dynamo probes find_spec("tensorflow")
            m.__dict__.update(attrs)
            sys.modules[name] = m
        return sys.modules[name]

    stub("tensorflow", random=types.SimpleNamespace(set_seed=lambda s: None))       # utils.py:11,42
    stub("warmup_scheduler", GradualWarmupScheduler=object)                          # utils.py:18
    stub("seaborn")
    mp = stub("matplotlib")
    stub("matplotlib.pyplot")
    mp.pyplot = sys.modules["matplotlib.pyplot"]
    stub("kneed", KneeLocator=object)
    saved_path, saved_argv = list(sys.path), list(sys.argv)
    shadow = {k: sys.modules.pop(k) for k in ("interpolation_layer", "rbf", "dec", "utils", "info") if k in sys.modules}
    sys.path.insert(0, DEST)
    sys.argv = ["x"]
    try:
        import interpolation_layer, dec, rbf                                          # noqa: E401
        ns = types.SimpleNamespace(interpolation_layer=interpolation_layer, dec=dec, rbf=rbf, path=DEST)
        if p2:
            import p2_clustering_optK
            ns.p2 = p2_clustering_optK
            sys.modules.pop("p2_clustering_optK", None)
            sys.modules.pop("internal_eval", None)
        for k in ("interpolation_layer", "rbf", "dec", "utils", "info"):              # leave no shadow behind
            sys.modules.pop(k, None)
        return ns
    finally:
        sys.path[:] = saved_path
        sys.argv[:] = saved_argv
        sys.modules.update(shadow)


def import_net_module(name="pretrain_interp", b200=False):
    """The staged reference's `pretrain_interp` / `clustering_interp` module.  b200=False: on the reference's own
    operators (the CPU arm).  b200=True: the same unmodified file on top of the B200 mirrors (dropin.install +
    dropin.patch_lstm) - what a user of the reference runs after switching.  Returns None when nothing is staged."""
    import importlib
    import importlib.machinery
    import types
    if not os.path.exists(os.path.join(DEST, name + ".py")):
        return None

    def stub(modname, **attrs):
        if modname not in sys.modules:
            m = types.ModuleType(modname)
            m.__spec__ = importlib.machinery.ModuleSpec(modname, None)
            m.__dict__.update(attrs)
            sys.modules[modname] = m

    stub("tensorflow", random=types.SimpleNamespace(set_seed=lambda s: None))
    stub("warmup_scheduler", GradualWarmupScheduler=object)
    for k in (name, "interpolation_layer", "rbf", "dec", "utils", "info"):
        sys.modules.pop(k, None)
    sys.path.insert(0, DEST)
    try:
        if b200:
            from deep_interpolation_clustering_b200 import dropin
            dropin.install()
            mod = importlib.import_module(name)
            dropin.patch_lstm(mod)
        else:
            mod = importlib.import_module(name)
        return mod
    finally:
        sys.path.remove(DEST)
        for k in (name, "interpolation_layer", "rbf", "dec", "utils", "info"):
            sys.modules.pop(k, None)


if __name__ == "__main__":
    if stage() is None:
        sys.exit(f"{REFERENCE} is not present: nothing staged")
