"""No GPU: libogn_b200.so loads and exports every symbol that include/*.h declares; nothing is computed."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for f in ("ogn_hierarchy_api.h", "ogn_b200.h", "ogn_kernels_api.h", "ogn_image_encoder_api.h"):
        text = open(os.path.join(ROOT, "include", f)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(ogn_[a-z0-9_]+)\s*\(", text))
        if f in ("ogn_hierarchy_api.h", "ogn_image_encoder_api.h"):
            names |= {"ogn_" + m for m in re.findall(r"P##([a-z0-9_]+)\s*\(", text)}
    return {n for n in names if not n.endswith("_fn")}


def test_library_exports_every_declared_symbol(product):
    lib = product.cdll()
    names = declared_symbols()
    assert len(names) > 45
    missing = sorted(n for n in names if not hasattr(lib, n))
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_flat_binding_resolves(product):
    from ogmaneo2_b200.flat_api import FLAT_API_NAMES
    from ogmaneo2_b200.hierarchy import PRODUCT_API_NAMES, _bind_product
    flat = _bind_product(product.lib())
    for n in FLAT_API_NAMES + PRODUCT_API_NAMES:
        assert callable(getattr(flat, n))


def test_no_cpu_fallback_without_device(product):
    """On a box without a usable sm_100 device, creating a hierarchy must fail loudly (never fall back to the CPU)."""
    lib = product.cdll()
    lib.ogn_device_count.restype = int
    if lib.ogn_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        product.Hierarchy([(4, 4, 16)], [product.InputType.prediction], [product.LayerDesc()])


def test_product_does_not_reference_the_oracle():
    """Nothing under ogmaneo2_b200/ or include/ may import, link or load oracle/ (it is test infrastructure)."""
    bad = []
    for base in ("ogmaneo2_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "_build" in dp or "__pycache__" in dp:
                continue
            for f in files:
                if not f.endswith((".py", ".h", ".cpp", ".cu")):
                    continue
                text = open(os.path.join(dp, f)).read()
                code = "\n".join(l for l in text.splitlines() if not l.strip().startswith(("//", "#", "*", "/*")))
                if re.search(r"^\s*(from|import)\s+oracle\b|sph_oracle|libogmaneo_ref|#include\s*[<\"][^>\"]*oracle", code, flags=re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
