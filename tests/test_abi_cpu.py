"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the packed weight layout covers the reference state_dict exactly, and the Python
classes keep the reference's names / error behaviour.  No kernel is launched here."""
import os
import re

import pytest
import torch

from elg_b200 import _lib, engine
from elg_b200.synth import DEFAULT_MODEL_PARAMS, state_dict_spec, synthetic_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "elg_b200.h")).read()
    declared = set(re.findall(r"\b(elg_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(_lib.lib, name) is not None
    assert _lib.lib.elg_abi_version() == _lib.ABI_VERSION


@pytest.mark.parametrize("problem", ["cvrp", "tsp"])
def test_weight_layout_covers_state_dict(problem):
    mp = DEFAULT_MODEL_PARAMS[problem]
    desc = engine.make_desc(problem, mp)
    slots, total = engine.weight_slots(problem, desc)
    spec = {k: shape for k, shape, _ in state_dict_spec(problem)}
    assert set(slots) == set(spec)
    used = sorted((off, off + int(torch.Size(spec[k]).numel())) for k, off in slots.items())
    for (a0, a1), (b0, b1) in zip(used, used[1:]):
        assert a1 <= b0, "overlapping weight slots"
    assert used[-1][1] <= total
    assert all(off % 4 == 0 for off in slots.values()), "slots must be 16-byte aligned for float4 loads"
    assert total == sum(int(torch.Size(s).numel()) for s in spec.values())   # released config needs no padding
    assert int(_lib.lib.elg_derived_floats(desc)) > 0


def test_unsupported_configs_are_rejected_loudly():
    mp = dict(DEFAULT_MODEL_PARAMS["cvrp"], embedding_dim=64)
    L = _lib.WeightLayout()
    assert _lib.lib.elg_weight_layout(engine.make_desc("cvrp", mp), L) == -2      # ELG_EUNSUPPORTED
    assert b"embedding_dim" in _lib.lib.elg_last_error()
    with pytest.raises(_lib.ElgError):
        engine.make_desc("cvrp", dict(DEFAULT_MODEL_PARAMS["cvrp"], ensemble_size=2))
    with pytest.raises(_lib.ElgError):
        engine.make_desc("tsp", dict(DEFAULT_MODEL_PARAMS["tsp"], euclidean=True))


@pytest.mark.parametrize("problem", ["cvrp", "tsp"])
def test_drop_in_classes_keep_reference_names(problem):
    if problem == "cvrp":
        from elg_b200.cvrp import CVRPEnv as Env, CVRPModel as Model, rollout
    else:
        from elg_b200.tsp import TSPEnv as Env, TSPModel as Model, rollout
    model = Model(**DEFAULT_MODEL_PARAMS[problem])
    assert not any("local" in k for k in model.state_dict())
    model.decoder.add_local_policy("cpu")              # must precede load_state_dict, as in the reference
    sd = synthetic_state_dict(problem)
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(sd)
    for name in ("pre_forward", "one_step_rollout", "encoder", "decoder", "encoded_nodes"):
        assert hasattr(model, name)
    for name in ("load_random_problems", "reset", "pre_step", "step", "selected_node_list"):
        assert hasattr(Env, name)
    assert callable(rollout)
    with pytest.raises(_lib.ElgError):                  # no CPU path: fail loudly
        Env(4, "cpu")
    with pytest.raises(_lib.ElgError):
        engine.ModelHandle(problem, DEFAULT_MODEL_PARAMS[problem], sd, "cpu")


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "elg_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "elg_oracle" not in src, f


def test_training_entry_points_validate_arguments_without_a_gpu():
    """Argument errors are reported as ELG_E* codes before any CUDA call (no compute on the CPU box)."""
    import ctypes as C
    from elg_b200 import _lib, engine
    from elg_b200.synth import DEFAULT_MODEL_PARAMS
    lib = _lib.lib
    desc = engine.make_desc("cvrp", dict(DEFAULT_MODEL_PARAMS["cvrp"]))
    null = C.c_void_p(0)
    assert lib.elg_train_saved_bytes(desc, 64, 101) == (6 * (8 * 128 + 512) + 128) * 64 * 101 * 4
    assert lib.elg_train_workspace_bytes(desc, 64, 100, 101, 204, 32) > lib.elg_train_workspace_bytes(desc, 64, 100, 101, 204, 1) > 0
    assert lib.elg_train_workspace_bytes(desc, 0, 100, 101, 204, 1) == 0
    out = (C.c_int64 * 8)()
    assert lib.elg_train_workspace_layout(desc, 4, 20, 21, 44, out) == 0 and list(out)[:4] == sorted(list(out)[:4]) and out[0] == 0
    t = _lib.Tables()
    assert lib.elg_encode_train(desc, null, null, t, 1, 21, null, 0, null) == -1           # ELG_EINVAL
    assert lib.elg_reinforce_backward(desc, null, null, t, null, 1, 20, 21, null, 44, 40, null, null, 1, null, null, null, 0, null) == -1
    assert b"NULL" in lib.elg_last_error()
    assert lib.elg_adam_step(null, null, null, null, 10, 1, 1e-4, 0.9, 0.999, 1e-8, 1e-6, 1.0, null) == -1
    assert lib.elg_generate_problems(1, 5, 4, 20, 3, 0.2, 0.8, 0.07, 30.0, 1, null, null, null, null) == -1
    bad = engine.make_desc("cvrp", dict(DEFAULT_MODEL_PARAMS["cvrp"]))
    bad.emb = 64
    assert lib.elg_train_saved_bytes(bad, 4, 21) == 0
