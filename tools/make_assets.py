#!/usr/bin/env python
"""Regenerate rsr_mjx_b200/assets/*.xml from the reference's MJCF models.

The reference ships its three Airbot models as MJCF with nested default classes
and render-only content (test/sf.xml, ppo_train/airbot_training/cube.xml,
ppo_train/airbot_training/T_shape.xml).  The GPU box has no /root/reference, so
the models travel as *flattened* MJCF: defaults resolved into explicit
attributes, attributes sorted, visuals dropped.  tests/test_mjcf.py checks (when
/root/reference is mounted) that the flattened file compiles to the identical
model as the original.
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from rsr_mjx_b200 import mjcf

REF = os.environ.get("RSR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(__file__), "..", "rsr_mjx_b200", "assets")
MODELS = {
    "airbot_sf.xml": "test/sf.xml",
    "airbot_cube.xml": "ppo_train/airbot_training/cube.xml",
    "airbot_T.xml": "ppo_train/airbot_training/T_shape.xml",
}
if __name__ == "__main__":
    for dst, src in MODELS.items():
        mjcf.write_flat_mjcf(os.path.join(REF, src), os.path.join(OUT, dst),
                             header=f"flattened by tools/make_assets.py from reference {src}")
        print("wrote", dst)
