"""Loader for the lane-1 host build of the kernel source (TEST HARNESS ONLY, see lane1.cpp)."""
import ctypes as C
import os
import subprocess

from mj_grasp_sim_b200 import lib as mlib

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(f64=False):
    so = os.path.join(_HERE, "liblane1_f64.so" if f64 else "liblane1_f32.so")
    srcs = [os.path.join(_HERE, "lane1.cpp")] + mlib._sources()
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared"] + (["-DMGS_REAL_DOUBLE"] if f64 else []) + [
            "-x", "c++", "-o", so, os.path.join(_HERE, "lane1.cpp")]
        subprocess.check_call(cmd)
    return so


def sim(model, f64=False):
    so = build(f64)
    if so not in _LIBS:
        _LIBS[so] = mlib.bind(C.CDLL(so), prefix="l1_")
    return mlib.BatchSim(model, lib=_LIBS[so], prefix="l1_")
