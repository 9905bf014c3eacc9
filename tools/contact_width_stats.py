#!/usr/bin/env python
"""Distribution of the Jacobian base-row storage a substep needs: sum over active contacts of the dof columns
the contact touches (tree dofs of its two bodies).  Sizes the variable-width row pool."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from rsr_mjx_b200 import prng, _lib
from rsr_mjx_b200.envs import AirbotPlayBase
kind = sys.argv[1] if len(sys.argv) > 1 else "sf"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
env = AirbotPlayBase(kind, num_envs=N, episode_length=1200)
m = env.model
# columns per geom: dofs of the kinematic tree of the geom's body (0 for static bodies)
body_tree_dofs = np.zeros(m.nbody, int)
root = np.arange(m.nbody)
for b in range(1, m.nbody):
    p = int(m.body_parentid[b])
    root[b] = root[p] if (p != 0 and (m.body_dofnum[p] > 0 or root[p] != p)) else b
moving = np.zeros(m.nbody, bool)
for b in range(1, m.nbody):
    moving[b] = m.body_dofnum[b] > 0 or moving[int(m.body_parentid[b])]
for b in range(m.nbody):
    if moving[b]:
        r = b
        while int(m.body_parentid[r]) != 0 and moving[int(m.body_parentid[r])]:
            r = int(m.body_parentid[r])
        root[b] = r
tree_dofs = {r: int(sum(m.body_dofnum[b] for b in range(m.nbody) if moving[b] and root[b] == r)) for r in set(root[moving])}
geom_cols = np.array([tree_dofs[root[b]] if moving[b] else 0 for b in m.geom_bodyid])
print("tree dofs", tree_dofs, "geom cols", geom_cols.tolist())
MAXC = _lib.lib().rsrx_max_contacts()
st = env.reset(prng.split(prng.PRNGKey(0), N))
gen = torch.Generator("cuda").manual_seed(0)
gc = torch.as_tensor(geom_cols, device="cuda")
widths = []
for t in range(T):
    env.step(st, torch.rand(N, 5, device="cuda", generator=gen) * 2 - 1)
    if t % 10 == 9:
        d = env.physics_step_debug(st._buf["data"].clone())
        nc = d[:, 480].long()
        cg = d[:, 483 + 4 * MAXC: 483 + 5 * MAXC].long().clamp(0, 64 * len(geom_cols) - 1)
        cg = torch.where(torch.arange(MAXC, device="cuda")[None] < nc[:, None], cg, torch.zeros_like(cg))
        w = gc[(cg // 64).clamp(max=len(geom_cols) - 1)] + gc[(cg % 64).clamp(max=len(geom_cols) - 1)]
        w = w * (torch.arange(MAXC, device="cuda")[None] < nc[:, None])
        widths.append(w.sum(1).cpu().numpy())
w = np.concatenate(widths)
print("columns per substep: mean %.1f" % w.mean(), {q: int(np.quantile(w, q)) for q in (0.5, 0.9, 0.99, 0.999, 0.9999, 1.0)})
print("pool floats (4 rows per contact) at p99.99 / max:", 4 * int(np.quantile(w, 0.9999)), 4 * int(w.max()))
