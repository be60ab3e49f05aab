"""Loads the byte-compiled, UNMODIFIED reference script from oracle/_ref/ (see oracle/build_ref.py).
TEST / MEASUREMENT INFRASTRUCTURE ONLY (tests/, bench.py's cpu_baseline and --impl reference legs)."""
import importlib.machinery
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def load(name="retrieval_data_annotation"):
    """The reference module (its `if __name__ == '__main__'` block does not run), or None when oracle/_ref/ has not
    been built (the reference tree was not mounted at build time)."""
    if name in _cache:
        return _cache[name]
    path = os.path.join(HERE, "_ref", name + ".bytecode")
    mod = None
    if os.path.exists(path):
        try:
            loader = importlib.machinery.SourcelessFileLoader("r4d_reference_" + name, path)
            spec = importlib.util.spec_from_loader(loader.name, loader)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
        except Exception:   # e.g. bytecode of another interpreter version
            mod = None
    _cache[name] = mod
    return mod
