"""On-disk RSR datasets: the six comma-separated text tables a reference user already has.

Mirrors `test/rsr_policy_training.py:49-205` (file names, truncation law, validation order and error types) and the
loader of `test/rsr_env_params_tuning.py:53-71`, so the same `data/` directory works unchanged:

    real_obs.txt, real_action.txt        real robot: observations s_0..s_T and actions a_0..a_{T-1}
    past_sim_obs.txt                     the same action sequence replayed in the previous simulator
    current_sim_obs.txt                  ... and in the simulator with the tuned parameters
    obs.txt, actions.txt                 sim rollouts (only validated, like upstream)

One row per time step, one column per feature, `,` separated, blank lines ignored.  Everything here is host-side
numpy; the arrays go to `rsr_pipeline.policy_params_training` / `env_params_tuning`, which move them to the GPU.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Mapping, Tuple

import numpy as np

REQUIRED_DATA_FILES = (
    "real_obs.txt",
    "real_action.txt",
    "past_sim_obs.txt",
    "current_sim_obs.txt",
    "obs.txt",
    "actions.txt",
)
MAX_TRANSITIONS = 50  # rsr_policy_training.py:61


def require_data_file(data_dir, filename: str) -> Path:
    """rsr_policy_training.py:69-76"""
    path = Path(data_dir) / filename
    if not path.is_file():
        raise FileNotFoundError(f"Required dataset file not found: {path}. "
                                f"Expected files: {', '.join(REQUIRED_DATA_FILES)}")
    return path


def load_numeric_table(path) -> np.ndarray:
    """A [rows, features] float64 table (rsr_policy_training.py:79-85; blank lines skipped like
    rsr_env_params_tuning.py:60-70).  A single row stays rank 2; an empty file is an error."""
    path = Path(path)
    rows = []
    with open(path, "r") as f:
        for line_no, line in enumerate(f, 1):
            text = line.strip()
            if not text:
                continue
            try:
                rows.append([float(tok) for tok in text.split(",")])
            except ValueError as e:
                raise ValueError(f"{path.name}:{line_no}: {e}") from None
    if not rows:
        raise ValueError(f"{path.name} is empty.")
    width = len(rows[0])
    for i, r in enumerate(rows):
        if len(r) != width:
            raise ValueError(f"{path.name}: row {i + 1} has {len(r)} columns, expected {width}.")
    return np.asarray(rows, dtype=np.float64)


def load_transition_triplet(obs_path, action_path, max_transitions: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(s_t, a_t, s_{t+1}) with a shared transition count = min(len(obs) - 1, len(actions), max_transitions)
    (rsr_policy_training.py:88-113)."""
    observations = load_numeric_table(obs_path)
    actions = load_numeric_table(action_path)
    count = min(len(observations) - 1, len(actions), max_transitions)
    if count <= 0:
        raise ValueError(f"Not enough aligned transitions in {Path(obs_path).name} and {Path(action_path).name}. "
                         "Need at least 2 observations and 1 action.")
    return observations[:count], actions[:count], observations[1:count + 1]


def _validate_observation_sequence(path, transition_count: int) -> np.ndarray:
    observations = load_numeric_table(path)
    need = transition_count + 1
    if len(observations) < need:
        raise ValueError(f"{Path(path).name} needs at least {need} rows for {transition_count} transitions, "
                         f"found {len(observations)}.")
    return observations


def _validate_action_sequence(path, transition_count: int) -> np.ndarray:
    actions = load_numeric_table(path)
    if len(actions) < transition_count:
        raise ValueError(f"{Path(path).name} needs at least {transition_count} rows, found {len(actions)}.")
    return actions


def _validate_feature_width(arrays: Mapping[str, np.ndarray], expected_width: int, label: str) -> None:
    for name, array in arrays.items():
        if array.shape[1] != expected_width:
            raise ValueError(f"{name} must have {expected_width} {label} features, found shape {array.shape}.")


def load_rsr_datasets(data_dir, max_transitions: int = MAX_TRANSITIONS, verbose: bool = False):
    """Loads and validates everything `policy_params_training` needs (rsr_policy_training.py:149-205).

    Returns float32 arrays `(past_states, past_actions, past_next_states_real, past_next_states_sim,
    current_next_states_sim)`, each with `transition_count` rows."""
    paths: Dict[str, Path] = {name: require_data_file(data_dir, name) for name in REQUIRED_DATA_FILES}
    past_states, past_actions, past_next_states_real = load_transition_triplet(
        paths["real_obs.txt"], paths["real_action.txt"], max_transitions)
    count, obs_dim, action_dim = past_states.shape[0], past_states.shape[1], past_actions.shape[1]
    past_sim_obs = _validate_observation_sequence(paths["past_sim_obs.txt"], count)
    current_sim_obs = _validate_observation_sequence(paths["current_sim_obs.txt"], count)
    sim_obs = _validate_observation_sequence(paths["obs.txt"], count)
    sim_actions = _validate_action_sequence(paths["actions.txt"], count)
    _validate_feature_width({"real_obs.txt": load_numeric_table(paths["real_obs.txt"]), "past_sim_obs.txt": past_sim_obs,
                             "current_sim_obs.txt": current_sim_obs, "obs.txt": sim_obs}, obs_dim, "observation")
    _validate_feature_width({"real_action.txt": load_numeric_table(paths["real_action.txt"]), "actions.txt": sim_actions},
                            action_dim, "action")
    if verbose:
        print("====== RSR dataset summary ======")
        print(f"data_dir: {data_dir}")
        print(f"transitions: {count}")
        for name in REQUIRED_DATA_FILES:
            print(f"{name}: {paths[name]}")
    f32 = np.float32
    return (past_states.astype(f32), past_actions.astype(f32), past_next_states_real.astype(f32),
            past_sim_obs[1:count + 1].astype(f32), current_sim_obs[1:count + 1].astype(f32))


def load_tuning_samples(real_obs_path, real_action_path, n: int = 15, index: int = 0):
    """The friction system-ID samples of rsr_env_params_tuning.py:75-104: `n` consecutive real transitions starting at
    `index`.  Returns float32 `(sampled_obs, sampled_actions, sampled_next_obs_true)`."""
    real_obs = load_numeric_table(real_obs_path)
    actions = load_numeric_table(real_action_path)
    obs = real_obs[index:index + n]
    act = actions[index:index + n]
    nxt = real_obs[1 + index:1 + index + n]
    m = min(len(obs), len(act), len(nxt))
    if m <= 0:
        raise ValueError(f"no aligned transitions at index {index} in {Path(real_obs_path).name} / "
                         f"{Path(real_action_path).name}")
    f32 = np.float32
    return obs[:m].astype(f32), act[:m].astype(f32), nxt[:m].astype(f32)


def write_numeric_table(path, table) -> None:
    """Inverse of `load_numeric_table` (`%.9g`: float32 round-trips exactly)."""
    table = np.asarray(table)
    if table.ndim != 2:
        raise ValueError(f"table must be rank 2, got shape {table.shape}")
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    np.savetxt(path, table, delimiter=",", fmt="%.9g")


def write_rsr_datasets(data_dir, real_obs, real_action, past_sim_obs, current_sim_obs, obs, actions) -> None:
    """Writes the six tables under `data_dir` (synthetic data for benchmarks/tests)."""
    tables = dict(zip(REQUIRED_DATA_FILES, (real_obs, real_action, past_sim_obs, current_sim_obs, obs, actions)))
    for name, table in tables.items():
        write_numeric_table(Path(data_dir) / name, table)


def load_dataset_from_path(path):
    """`(states, actions, next_states)` from an `.npz` with those three arrays (RSR/dataset_processor.py:10-14)."""
    with np.load(Path(path), allow_pickle=False) as data:
        missing = [k for k in ("states", "actions", "next_states") if k not in data]
        if missing:
            raise KeyError(f"{path}: missing arrays {missing}")
        return np.array(data["states"]), np.array(data["actions"]), np.array(data["next_states"])


def concatenate_states(states):
    """Stacks the observations of a list of env `State`s row-wise (RSR/dataset_processor.py:45-53); accepts our batched
    `State` (obs [N, obs_size]) as well as plain arrays."""
    rows = []
    for s in states:
        obs = getattr(s, "obs", s)
        obs = obs.detach().cpu().numpy() if hasattr(obs, "detach") else np.asarray(obs)
        rows.append(obs.reshape(-1, obs.shape[-1]))
    if not rows:
        raise ValueError("no states to concatenate")
    return np.vstack(rows)
