"""Seeded random hierarchy shapes (sizes, radii, temporal horizons, tick rates, several inputs) shared by the CPU test
that pins the oracle port against the compiled reference and the GPU test that checks the CUDA path on the same shapes."""
import numpy as np

from ogmaneo2_b200.flat_api import InputType, LayerDesc


def random_cfg(seed: int) -> dict:
    rng = np.random.RandomState(1000 + seed)
    zs = [3, 4, 5, 8, 12, 16, 20, 32, 33]
    n_in = int(rng.randint(1, 3))
    sizes = [(int(rng.randint(1, 8)), int(rng.randint(1, 8)), int(rng.choice(zs))) for _ in range(n_in)]
    types = [InputType.prediction if rng.rand() < 0.75 else InputType.none for _ in range(n_in)]
    if InputType.prediction not in types:
        types[0] = InputType.prediction
    lds = []
    for _ in range(int(rng.randint(1, 4))):
        tpu = int(rng.randint(1, 4))
        lds.append(LayerDesc(hiddenSize=(int(rng.randint(1, 8)), int(rng.randint(1, 8)), int(rng.choice(zs))),
                             ffRadius=int(rng.randint(0, 4)), pRadius=int(rng.randint(0, 4)), ticksPerUpdate=tpu,
                             temporalHorizon=tpu + int(rng.randint(0, 3))))
    return dict(name=f"rand{seed}", inputSizes=sizes, inputTypes=types, layerDescs=lds, streams=["iid"] * n_in, seed=77 + seed)
