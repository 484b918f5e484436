"""Golden vectors of the visual-token construction from the REFERENCE's own code (build container only).

    python tests/golden/make_ref_token_goldens.py        # writes tests/golden/ref_tokens.npz

The per-scene loop lives inside ``SIG3D.forward`` (situation3d/models/sqa_module.py:297-315); the module cannot be
imported (MinkowskiEngine, dataset-scanning config), so the ``for`` statement itself is taken out of the parsed
source and executed unmodified against stand-ins for the three names it reads: the decomposed coordinate /
feature lists, ``scene_feat_original.tensor_stride`` and ``CONF.OPENSCENE``.  Only inputs and outputs are stored.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/situation3d/models/sqa_module.py"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def reference_loop():
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SIG3D")
    fwd = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward")
    loops = [n for n in ast.walk(fwd) if isinstance(n, ast.For) and "batch_idx" in ast.dump(n.target)]
    assert len(loops) == 1
    return compile(ast.Module(body=[loops[0]], type_ignores=[]), REF, "exec")


def make_scene(g, m, extent, c):
    """Voxel coordinates at tensor stride 16 (multiples of 16, negative values included, several z per column)."""
    xy = torch.randint(-extent, extent, (m, 2), generator=g) * 16
    z = torch.randint(0, 6, (m, 1), generator=g) * 16
    coords = torch.cat([xy, z], dim=1).to(torch.int32)
    coords = torch.unique(coords, dim=0)                       # a sparse tensor has no duplicate voxels
    coords = coords[torch.randperm(coords.shape[0], generator=g)]
    return coords, torch.randn(coords.shape[0], c, generator=g)


def main():
    code = reference_loop()
    g = torch.Generator().manual_seed(777)
    out = {}
    cases = {"a": [(900, 14), (300, 6), (2500, 30)],      # more / fewer than 256 columns, a large scene
             "b": [(40, 3), (5000, 40)]}
    for name, scenes in cases.items():
        data = [make_scene(g, m, e, 64 if name == "a" else 32) for m, e in scenes]
        coords_list, feats_list = [d[0] for d in data], [d[1] for d in data]
        conf = types.SimpleNamespace(OPENSCENE=types.SimpleNamespace(num_points=256, voxel_size=0.02))
        ns = {"torch": torch, "CONF": conf, "list_of_coords": coords_list, "list_of_featurs": feats_list,
              "scene_feat_original": types.SimpleNamespace(tensor_stride=[16, 16, 16]), "scene_feat": [], "scene_positions": []}
        torch.manual_seed(1234)
        exec(code, ns)
        for i, (c, f) in enumerate(zip(coords_list, feats_list)):
            out["%s_coords%d" % (name, i)] = c.numpy()
            out["%s_feats%d" % (name, i)] = f.numpy()
        out[name + "_scene_feat"] = torch.cat(ns["scene_feat"], dim=0).numpy()
        out[name + "_scene_positions"] = torch.cat(ns["scene_positions"], dim=0).numpy()
        out[name + "_nscenes"] = np.array(len(scenes))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_tokens.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if "scene" in k})


if __name__ == "__main__":
    main()
