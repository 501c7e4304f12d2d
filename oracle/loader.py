"""oracle/loader.py — TEST INFRASTRUCTURE ONLY: locates / builds / loads the CPU checkers.

    port()  -> FlatLib for oracle/libsph_oracle.so  (prefix oport_), built with `make -C oracle port` if missing
    ref()   -> FlatLib for oracle/_ref/libogmaneo_ref.so (prefix oref_) or None. It is only (re)built where the
               reference sources are present (/root/reference, i.e. the build container); on the GPU box the
               prebuilt .so that travelled with the snapshot is used.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import subprocess

from ogmaneo2_b200.flat_api import FlatLib

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libsph_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libogmaneo_ref.so")
REF_SRC = os.environ.get("OGN_REF_DIR", "/root/reference")

_cache = {}


def _digest(sources):
    import hashlib
    h = hashlib.sha256()
    for s in sources:
        if os.path.exists(s):
            with open(s, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def _stale(target, sources):
    """Content-hash staleness (mtimes do not survive the snapshot that travels to the GPU box)."""
    stamp = target + ".sha"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    with open(stamp) as f:
        return f.read().strip() != _digest(sources)


def _stamp(target, sources):
    with open(target + ".sha", "w") as f:
        f.write(_digest(sources))


def build(verbose=False):
    """Compile the port always (if stale) and oracle/_ref when the reference sources are present."""
    out = None if verbose else subprocess.DEVNULL
    srcs = [os.path.join(HERE, "sph_oracle.cpp"), os.path.join(HERE, "..", "include", "ogn_hierarchy_api.h"),
            os.path.join(HERE, "ienc_oracle.cpp"), os.path.join(HERE, "..", "include", "ogn_image_encoder_api.h")]
    if _stale(PORT_SO, srcs):
        subprocess.check_call(["make", "-B", "-C", HERE, "port"], stdout=out)
        _stamp(PORT_SO, srcs)
    if os.path.isdir(os.path.join(REF_SRC, "source", "ogmaneo")):
        rsrcs = [os.path.join(HERE, "ref_shim.cpp"), srcs[1], srcs[3]]
        if _stale(REF_SO, rsrcs):
            subprocess.check_call(["make", "-C", HERE, "ref", f"REF={REF_SRC}"], stdout=out)
            _stamp(REF_SO, rsrcs)


def port() -> FlatLib:
    if "port" not in _cache:
        build()
        _cache["port"] = FlatLib(ctypes.CDLL(PORT_SO), "oport_")
    return _cache["port"]


def ref():
    if "ref" not in _cache:
        try:
            build()
        except Exception:
            pass
        _cache["ref"] = FlatLib(ctypes.CDLL(REF_SO), "oref_") if os.path.exists(REF_SO) else None
    return _cache["ref"]


def best():
    """(FlatLib, kind): the compiled reference when available ("reference"), else the port ("port")."""
    r = ref()
    return (r, "reference") if r is not None else (port(), "port")
