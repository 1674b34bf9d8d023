"""The Lightning-free generation trainer (c4a0_b200/training.py) on CPU, modelled on the reference's
tests/c4a0_tests/nn_test.py and training_test.py (loss of self-labels is zero; val_loss equals the
loss computed by hand; saved model differs from its parent; artefacts round-trip)."""

import os

import numpy as np
import pytest
import torch

import c4a0_rust as R
import oracle
from c4a0_b200 import training as T
from c4a0_b200.nn import ConnectFourNet, ModelConfig

CFG = ModelConfig(n_residual_blocks=1, conv_filter_size=4, n_policy_layers=2, n_value_layers=2, lr_schedule={0: 2e-3, 10: 8e-4}, l2_reg=4e-4)


def _games(n_games=12, n_iter=10):
    reqs = [(i, 0, 0) for i in range(n_games)]
    out = oracle.self_play(reqs, 4, n_iter, 2.0, 0.01, evaluator="hash")
    games = []
    for r, ss in zip(reqs, out.samples):
        games.append(R.GameResult(R.GameMetadata(*r), [R.Sample(s.pos.mask, s.pos.value, list(s.policy), s.q_penalty, s.q_no_penalty) for s in ss]))
    return R.PlayGamesResult._from_results(games)


def test_lr_schedule():
    s = T.parse_lr_schedule([0, 2e-3, 10, 8e-4])
    assert s == {0: 2e-3, 10: 8e-4}
    assert [T.lr_for_generation(s, g) for g in (0, 1, 9, 10, 50)] == [2e-3, 2e-3, 2e-3, 8e-4, 8e-4]
    with pytest.raises(ValueError):
        T.parse_lr_schedule([0, 1e-3, 5])
    with pytest.raises(ValueError):
        T.parse_lr_schedule([0.5, 1e-3])


def test_loss_of_self_labels_is_zero_and_terms_add_up():
    torch.manual_seed(0)
    model = ConnectFourNet(CFG).eval()
    x = torch.zeros(8, 2, 6, 7)
    x[:, 0, 0, 3] = 1
    with torch.no_grad():
        logp, qp, qn = model(x)
        total, kl, a, b = T.loss_terms(model, x, logp.exp(), qp, qn)
    assert abs(float(total)) < 1e-5 and float(kl) > -1e-6
    with torch.no_grad():
        total, kl, a, b = T.loss_terms(model, x, torch.full((8, 7), 1 / 7), qp + 0.5, qn - 1.0)
    assert abs(float(a) - 0.25) < 1e-6 and abs(float(b) - 1.0) < 1e-6 and float(kl) > 0
    assert abs(float(total) - float(kl + a + b)) < 1e-6


def test_split_arrays_equals_sample_path_with_flip_augmentation():
    res = _games()
    (tp, tpol, tqp, tqn), (vp, vpol, vqp, vqn) = T.split_arrays(res, 0.8, 1337, augment=True)
    train, test = res.split_train_test(0.8, 1337)
    for samples, (pos, pol, qp, qn) in ((train, (tp, tpol, tqp, tqn)), (test, (vp, vpol, vqp, vqn))):
        both = samples + [s.flip_h() for s in samples]  # training.py:316-317
        assert len(both) == len(pos)
        for i in (0, len(samples) - 1, len(samples), len(both) - 1):
            p, l, a, b = both[i].to_numpy()
            assert np.array_equal(pos[i], p) and np.array_equal(pol[i], l) and qp[i] == a and qn[i] == b


def test_fit_improves_keeps_best_and_reports_manual_val_loss():
    res = _games(n_games=24)
    train, val = T.split_arrays(res)
    torch.manual_seed(3)
    parent = ConnectFourNet(CFG)
    before = {k: v.clone() for k, v in parent.state_dict().items()}
    best, val_loss, epochs = T.fit(parent, train, val, batch_size=64, lr=2e-3, l2_reg=4e-4, device=torch.device("cpu"), max_epochs=12, patience=3)
    assert all(torch.equal(v, before[k]) for k, v in parent.state_dict().items())  # the parent is not touched
    assert any(not torch.equal(v, before[k]) for k, v in best.state_dict().items() if v.dtype.is_floating_point)
    with torch.no_grad():
        manual = float(T.loss_terms(best.eval(), *[torch.from_numpy(a) for a in val])[0])
        start = float(T.loss_terms(parent.eval(), *[torch.from_numpy(a) for a in val])[0])
    assert abs(manual - val_loss) < 1e-5  # training_test.py: val_loss equals the manual loss
    assert 1 <= epochs <= 12 and np.isfinite(start)


def test_fit_learns_a_learnable_target():
    """Labels produced by a fixed teacher network are learnable: the validation loss must drop."""
    torch.manual_seed(5)
    teacher = ConnectFourNet(CFG).eval()
    rng = np.random.default_rng(0)
    cells = rng.integers(0, 3, size=(600, 42))
    x = np.zeros((600, 2, 42), np.float32)
    x[:, 0][cells == 1] = 1
    x[:, 1][cells == 2] = 1
    x = x.reshape(600, 2, 6, 7)
    with torch.no_grad():
        logp, qp, qn = teacher(torch.from_numpy(x))
    data = (x, logp.exp().numpy(), qp.numpy(), qn.numpy())
    train = tuple(a[:500] for a in data)
    val = tuple(a[500:] for a in data)
    torch.manual_seed(6)
    student = ConnectFourNet(CFG)
    with torch.no_grad():
        start = float(T.loss_terms(student.eval(), *[torch.from_numpy(a) for a in val])[0])
    best, val_loss, epochs = T.fit(student, train, val, batch_size=100, lr=2e-3, l2_reg=0.0, device=torch.device("cpu"), max_epochs=30, patience=30)
    assert val_loss < 0.7 * start, (start, val_loss)


def test_generation_directory_round_trip(tmp_path):
    base = str(tmp_path / "training")
    params = dict(n_mcts_iterations=4, c_exploration=6.6, c_ply_penalty=0.01, self_play_batch_size=8, training_batch_size=16)
    gen0 = T.TrainingGen.load_latest_with_default(base, CFG, **params)
    assert gen0.gen_n == 0 and os.path.exists(os.path.join(gen0.gen_folder(base), "model.pkl"))
    assert T.TrainingGen.load_latest_with_default(base, CFG, **params).created_at == gen0.created_at
    model = gen0.get_model(base)
    assert isinstance(model, ConnectFourNet) and gen0.get_games(base) is None
    from datetime import datetime

    gen1 = T.TrainingGen(created_at=datetime.now(), gen_n=1, parent=gen0.created_at, val_loss=1.25, **params)
    gen1.save_all(base, _games(4, 4), model)
    gens = T.TrainingGen.load_all(base)
    assert [g.gen_n for g in gens] == [1, 0] and gens[0].val_loss == 1.25
    assert len(gens[0].get_games(base).results) == 4
    assert sorted(os.listdir(gens[0].gen_folder(base))) == ["games.pkl", "metadata.json", "model.pkl", "model_state.pt"]


def test_model_pkl_written_by_the_reference_can_be_resumed(tmp_path):
    """model.pkl of a reference training directory names `c4a0.nn.ConnectFourNet`, a LightningModule
    with torchmetrics members (src/c4a0/nn.py:41-57, training.py:62-67).  Emulated here with throw-away modules
    `c4a0.nn`, `pytorch_lightning`, `torchmetrics` that exist only while the file is written."""
    import pickle
    import sys
    import types

    from c4a0_b200.nn import ConnectFourNet, ModelConfig
    from c4a0_b200.training import TrainingGen

    cfg = ModelConfig(n_residual_blocks=1, conv_filter_size=4, n_policy_layers=3, n_value_layers=2)
    torch.manual_seed(3)
    ours = ConnectFourNet(cfg)

    pl = types.ModuleType("pytorch_lightning")
    tm = types.ModuleType("torchmetrics")
    ref_pkg, ref_nn = types.ModuleType("c4a0"), types.ModuleType("c4a0.nn")

    class LightningModule(torch.nn.Module):
        pass

    class MeanMetric(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.register_buffer("mean_value", torch.zeros(()))

    from typing import Dict

    from pydantic import BaseModel

    class RefModelConfig(BaseModel):  # the reference's ModelConfig is a pydantic model too (nn.py:16-38)
        n_residual_blocks: int
        conv_filter_size: int
        n_policy_layers: int
        n_value_layers: int
        lr_schedule: Dict[int, float] = {0: 2e-3}
        l2_reg: float = 4e-4

    class RefNet(LightningModule):
        def __init__(self, c):
            super().__init__()
            self.config = RefModelConfig(**c.model_dump())
            self.conv, self.fc_policy, self.fc_value = ours.conv, ours.fc_policy, ours.fc_value
            self.policy_kl_div, self.value_mse = MeanMetric(), MeanMetric()
            self._trainer = None

    for mod, cls, name in ((pl, LightningModule, "LightningModule"), (tm, MeanMetric, "MeanMetric"),
                           (ref_nn, RefNet, "ConnectFourNet"), (ref_nn, RefModelConfig, "ModelConfig")):
        cls.__module__, cls.__qualname__, cls.__name__ = mod.__name__, name, name
        setattr(mod, name, cls)
    ref_nn.ResidualBlock = type(list(ours.conv.children())[1])
    saved = {k: sys.modules.get(k) for k in ("pytorch_lightning", "torchmetrics", "c4a0", "c4a0.nn")}
    sys.modules.update({"pytorch_lightning": pl, "torchmetrics": tm, "c4a0": ref_pkg, "c4a0.nn": ref_nn})
    try:
        gen = TrainingGen(created_at=__import__("datetime").datetime(2024, 1, 2, 3, 4, 5), gen_n=3, n_mcts_iterations=4,
                          c_exploration=1.0, c_ply_penalty=0.01, self_play_batch_size=4, training_batch_size=4)
        d = gen.gen_folder(str(tmp_path))
        os.makedirs(d)
        with open(os.path.join(d, "model.pkl"), "wb") as f:
            pickle.dump(RefNet(cfg), f)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert "pytorch_lightning" not in sys.modules and "c4a0.nn" not in sys.modules
    loaded = gen.get_model(str(tmp_path))
    assert type(loaded) is ConnectFourNet and loaded.config.n_policy_layers == 3 and loaded.config.conv_filter_size == 4
    for (ka, a), (kb, b) in zip(sorted(ours.state_dict().items()), sorted(loaded.state_dict().items())):
        assert ka == kb and torch.equal(a, b)
    x = (torch.rand(5, 2, 6, 7) < 0.3).float()
    for a, b in zip(ours.eval()(x), loaded(x)):
        assert torch.equal(a, b)


def test_generation_directory_also_holds_a_class_free_model_file(tmp_path):
    from c4a0_b200.nn import ConnectFourNet, ModelConfig
    from c4a0_b200.training import TrainingGen

    cfg = ModelConfig(n_residual_blocks=1, conv_filter_size=2, n_policy_layers=2, n_value_layers=2)
    gen = TrainingGen.load_latest_with_default(str(tmp_path), cfg, n_mcts_iterations=2, c_exploration=1.0, c_ply_penalty=0.01,
                                               self_play_batch_size=4, training_batch_size=4)
    blob = torch.load(os.path.join(gen.gen_folder(str(tmp_path)), "model_state.pt"), weights_only=False)
    m = ConnectFourNet(ModelConfig(**blob["config"]))
    m.load_state_dict(blob["state_dict"])
    os.remove(os.path.join(gen.gen_folder(str(tmp_path)), "model.pkl"))
    again = gen.get_model(str(tmp_path))  # falls back to the class-free file
    for a, b in zip(m.state_dict().values(), again.state_dict().values()):
        assert torch.equal(a, b)
