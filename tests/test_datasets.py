"""On-disk RSR tables: same rules as test/rsr_policy_training.py:69-205."""
import numpy as np
import pytest

from rsr_mjx_b200 import datasets as D


def _write_all(tmp_path, T=60, obs_dim=23, act_dim=5, seed=0, **override):
    g = np.random.default_rng(seed)
    tables = dict(real_obs=g.normal(size=(T + 1, obs_dim)), real_action=g.uniform(-1, 1, (T, act_dim)),
                  past_sim_obs=g.normal(size=(T + 1, obs_dim)), current_sim_obs=g.normal(size=(T + 1, obs_dim)),
                  obs=g.normal(size=(T + 1, obs_dim)), actions=g.uniform(-1, 1, (T, act_dim)))
    tables.update(override)
    D.write_rsr_datasets(tmp_path, tables["real_obs"].astype(np.float32), tables["real_action"].astype(np.float32),
                         tables["past_sim_obs"].astype(np.float32), tables["current_sim_obs"].astype(np.float32),
                         tables["obs"].astype(np.float32), tables["actions"].astype(np.float32))
    return {k: v.astype(np.float32) for k, v in tables.items()}


def test_round_trip_and_truncation(tmp_path):
    t = _write_all(tmp_path, T=60)
    S, A, S1r, S1p, S1c = D.load_rsr_datasets(tmp_path)  # MAX_TRANSITIONS = 50
    assert S.shape == (50, 23) and A.shape == (50, 5) and S.dtype == np.float32
    np.testing.assert_array_equal(S, t["real_obs"][:50])          # %.9g: float32 survives the text round trip
    np.testing.assert_array_equal(A, t["real_action"][:50])
    np.testing.assert_array_equal(S1r, t["real_obs"][1:51])
    np.testing.assert_array_equal(S1p, t["past_sim_obs"][1:51])
    np.testing.assert_array_equal(S1c, t["current_sim_obs"][1:51])
    # transition count = min(len(obs) - 1, len(actions), max_transitions)
    assert D.load_rsr_datasets(tmp_path, max_transitions=7)[0].shape[0] == 7
    t = _write_all(tmp_path, T=60, real_action=np.zeros((12, 5)))
    assert D.load_rsr_datasets(tmp_path)[0].shape[0] == 12


def test_missing_and_empty_files(tmp_path):
    _write_all(tmp_path)
    (tmp_path / "obs.txt").unlink()
    with pytest.raises(FileNotFoundError, match="obs.txt"):
        D.load_rsr_datasets(tmp_path)
    _write_all(tmp_path)
    (tmp_path / "actions.txt").write_text("\n\n")
    with pytest.raises(ValueError, match="actions.txt is empty"):
        D.load_rsr_datasets(tmp_path)


def test_validation_errors(tmp_path):
    _write_all(tmp_path, T=60, past_sim_obs=np.zeros((20, 23)))
    with pytest.raises(ValueError, match="past_sim_obs.txt needs at least 51 rows"):
        D.load_rsr_datasets(tmp_path)
    _write_all(tmp_path, T=60, actions=np.zeros((10, 5)))
    with pytest.raises(ValueError, match="actions.txt needs at least 50 rows"):
        D.load_rsr_datasets(tmp_path)
    _write_all(tmp_path, T=60, current_sim_obs=np.zeros((61, 22)))
    with pytest.raises(ValueError, match="current_sim_obs.txt must have 23 observation features"):
        D.load_rsr_datasets(tmp_path)
    _write_all(tmp_path, T=60, actions=np.zeros((60, 4)))
    with pytest.raises(ValueError, match="actions.txt must have 5 action features"):
        D.load_rsr_datasets(tmp_path)
    _write_all(tmp_path, T=60, real_obs=np.zeros((1, 23)))
    with pytest.raises(ValueError, match="Not enough aligned transitions"):
        D.load_rsr_datasets(tmp_path)


def test_table_parsing(tmp_path):
    p = tmp_path / "t.txt"
    p.write_text("1,2,3\n")
    assert D.load_numeric_table(p).shape == (1, 3)       # a single row stays rank 2
    p.write_text("\n1.5, 2e-3 ,3\n\n4,5,6\n   \n")
    np.testing.assert_array_equal(D.load_numeric_table(p), [[1.5, 2e-3, 3], [4, 5, 6]])
    p.write_text("1,2,3\n4,5\n")
    with pytest.raises(ValueError, match="row 2 has 2 columns"):
        D.load_numeric_table(p)
    p.write_text("1,x,3\n")
    with pytest.raises(ValueError, match="t.txt:1"):
        D.load_numeric_table(p)
    with pytest.raises(ValueError, match="rank 2"):
        D.write_numeric_table(p, np.zeros(3))


def test_tuning_samples(tmp_path):
    t = _write_all(tmp_path, T=40)
    o, a, n = D.load_tuning_samples(tmp_path / "real_obs.txt", tmp_path / "real_action.txt", n=15, index=3)
    np.testing.assert_array_equal(o, t["real_obs"][3:18])
    np.testing.assert_array_equal(a, t["real_action"][3:18])
    np.testing.assert_array_equal(n, t["real_obs"][4:19])
    o, a, n = D.load_tuning_samples(tmp_path / "real_obs.txt", tmp_path / "real_action.txt", n=15, index=35)
    assert len(o) == len(a) == len(n) == 5
    with pytest.raises(ValueError, match="no aligned transitions"):
        D.load_tuning_samples(tmp_path / "real_obs.txt", tmp_path / "real_action.txt", n=15, index=41)


def test_npz_and_state_helpers(tmp_path):
    g = np.random.default_rng(0)
    S, A, S1 = g.normal(size=(7, 23)), g.normal(size=(7, 5)), g.normal(size=(7, 23))
    np.savez(tmp_path / "d.npz", states=S, actions=A, next_states=S1)
    s, a, s1 = D.load_dataset_from_path(tmp_path / "d.npz")
    np.testing.assert_array_equal(s, S); np.testing.assert_array_equal(a, A); np.testing.assert_array_equal(s1, S1)
    np.savez(tmp_path / "bad.npz", states=S)
    with pytest.raises(KeyError, match="actions"):
        D.load_dataset_from_path(tmp_path / "bad.npz")

    class St:
        def __init__(self, o): self.obs = o
    out = D.concatenate_states([St(S[:3]), St(S[3]), S[4:]])
    np.testing.assert_array_equal(out, S)
    with pytest.raises(ValueError):
        D.concatenate_states([])
