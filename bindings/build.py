"""Builds bindings/kiss_icp_pybind*.so — the reference's `kiss_icp.pybind.kiss_icp_pybind` module re-bound over the C-ABI."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def module_path() -> str:
    return os.path.join(HERE, "kiss_icp_pybind" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(verbose: bool = False) -> str:
    import pybind11

    src, out = os.path.join(HERE, "kiss_icp_pybind.cpp"), module_path()
    lib = os.path.join(ROOT, "kiss-icp_b200", "libkiss_icp_b200.so")
    deps = [src, os.path.join(ROOT, "include", "kiss_icp_b200.h"), lib]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I" + pybind11.get_include(),
           "-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(ROOT, "include"), src,
           "-L" + os.path.dirname(lib), "-lkiss_icp_b200", "-Wl,-rpath,$ORIGIN/../kiss-icp_b200", "-o", out]
    if verbose:
        print(" ".join(cmd), flush=True)
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(cmd, check=True, env=env)
    return out


def load():
    """import the built module under its reference name"""
    import importlib.util

    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    spec = importlib.util.spec_from_file_location("kiss_icp_pybind", build())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose=True))
