"""The five BASELINE.json configs made concrete, plus scale knobs for test-sized variants.

Each config is a dict: inputSizes, inputTypes, layerDescs, stream kinds per input, seed. Unspecified LayerDesc
fields keep the reference defaults (Hierarchy.h:54-66).
"""
from .flat_api import InputType, LayerDesc


def c1_sine():
    """C1: 1x1x32 input, 4 layers of 4x4x16, radius 2 (CPU-runnable)."""
    return dict(name="C1", inputSizes=[(1, 1, 32)], inputTypes=[InputType.prediction],
                layerDescs=[LayerDesc(hiddenSize=(4, 4, 16), ffRadius=2, pRadius=2) for _ in range(4)],
                streams=["sine"], seed=1234)


def c2_video(n=32, layers=3):
    """C2: n x n x16 input, `layers` layers of n x n x32, radius 3, prediction on every layer."""
    return dict(name=f"C2@{n}", inputSizes=[(n, n, 16)], inputTypes=[InputType.prediction],
                layerDescs=[LayerDesc(hiddenSize=(n, n, 32), ffRadius=3, pRadius=3) for _ in range(layers)],
                streams=["noisy"], seed=1234)


def c3_actor(n=8):
    """C3: n x n x16 observation (no prediction) + 2x2x4 action input, 3 layers of n x n x16, Actor head."""
    return dict(name=f"C3@{n}", inputSizes=[(n, n, 16), (2, 2, 4)], inputTypes=[InputType.none, InputType.action],
                layerDescs=[LayerDesc(hiddenSize=(n, n, 16)) for _ in range(3)],
                streams=["noisy", "action"], seed=1234)


def c4_large(n=512, temporalHorizon=1):
    """C4: n x n x16 input, one layer of n x n x32 hidden columns, radius 4, temporalHorizon 1."""
    return dict(name=f"C4@{n}", inputSizes=[(n, n, 16)], inputTypes=[InputType.prediction],
                layerDescs=[LayerDesc(hiddenSize=(n, n, 32), ffRadius=4, pRadius=4, temporalHorizon=temporalHorizon)],
                streams=["noisy"], seed=1234)


def c5_unit(n=16, k=0):
    """C5 member k: n x n x16 input, 4 layers of n x n x16, radius 2; seed 1234+k, stream shifted by 17k."""
    return dict(name=f"C5@{n}#{k}", inputSizes=[(n, n, 16)], inputTypes=[InputType.prediction],
                layerDescs=[LayerDesc(hiddenSize=(n, n, 16)) for _ in range(4)],
                streams=["noisy"], seed=1234 + k, shift=17 * k, noise_seed=7 + k)
