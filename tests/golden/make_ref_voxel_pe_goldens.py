"""Golden vectors of the voxel-coordinate positional embedding from the REFERENCE's own code (build container only).

    python tests/golden/make_ref_voxel_pe_goldens.py        # writes tests/golden/ref_voxel_pe.npz

The code lives inside ``Blip2T5.forward`` (3DLLM_BLIP2-base/lavis/models/blip2_models/blip2_t5.py:104-118) and
``Blip2OPT.forward`` (blip2_opt.py:92-104); neither module can be imported here (LAVIS registry, T5 / OPT
checkpoints, the un-vendored ``positional_encodings`` package).  The statements from ``pc = samples["pc"].long()``
to the assignment of ``pc_embeds`` are taken out of the parsed source and executed unmodified on the CPU against
stand-ins for the names they read: ``samples``, ``pc_embeds`` and ``self.pos_embedding``.  ``Tensor.cuda()`` is the
identity while they run (the reference moves its CPU result to the GPU before the add; the values do not change).
Only inputs and outputs are stored.  The table is an input: the first case uses a random one, the second the first
64 rows of the sinusoid table as this repo builds it (situation3d_b200.voxel_pe.sinusoid_table).
"""
import ast
import os
import sys
import types

import numpy as np
import torch

REF_DIR = "/root/reference/3DLLM_BLIP2-base/lavis/models/blip2_models/"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def _flatten(stmts):
    for s in stmts:
        if isinstance(s, ast.With):
            yield from _flatten(s.body)
        else:
            yield s


def _assigns(stmt, name):
    return isinstance(stmt, ast.Assign) and any(isinstance(t, ast.Name) and t.id == name for t in stmt.targets)


def reference_statements(fname, cls_name):
    path = REF_DIR + fname
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name)
    fwd = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward")
    flat = list(_flatten(fwd.body))
    first = next(i for i, s in enumerate(flat) if _assigns(s, "pc"))
    loop = next(i for i, s in enumerate(flat) if isinstance(s, ast.For) and i > first)
    last = next(i for i, s in enumerate(flat) if i > loop and _assigns(s, "pc_embeds"))
    body = flat[first:last + 1]
    print(fname, "lines", body[0].lineno, "-", body[-1].end_lineno)
    return compile(ast.Module(body=body, type_ignores=[]), path, "exec")


def run(code, pc_feat, pc, table):
    ns = {"torch": torch, "samples": {"pc": pc, "pc_feat": pc_feat}, "pc_embeds": pc_feat.clone(),
          "self": types.SimpleNamespace(pos_embedding=table)}
    exec(code, ns)
    return ns["pc_embeds"], ns["all_pcs"]


def main():
    from situation3d_b200.voxel_pe import sinusoid_table
    t5 = reference_statements("blip2_t5.py", "Blip2T5")
    opt = reference_statements("blip2_opt.py", "Blip2OPT")
    torch.Tensor.cuda = lambda self, *a, **k: self
    g = torch.Generator().manual_seed(4242)
    out = {}
    cases = {"rand": (torch.randn(40, 469, generator=g), 2, 24, -40, 40),       # negative indices wrap
             "sin": (sinusoid_table(256, 469, "concat")[:64].clone(), 3, 16, 0, 64)}
    for name, (table, B, P, lo, hi) in cases.items():
        pc_feat = torch.randn(B, P, 1408, generator=g)
        # float coordinates with a fractional part, as voxelised clouds are stored: .long() truncates toward zero
        pc = torch.randint(lo, hi, (B, P, 3), generator=g).float()
        pc = pc + torch.where(pc >= 0, 1.0, -1.0) * torch.rand(B, P, 3, generator=g) * 0.9
        pc = torch.where(pc.long() < lo, pc.long().float(), pc)
        add, all_pcs = run(t5, pc_feat, pc, table)
        cat, all_pcs2 = run(opt, pc_feat, pc, table)
        assert torch.equal(all_pcs, all_pcs2) and torch.equal(cat, torch.cat([pc_feat, all_pcs], 1))
        out.update({name + "_table": table.numpy(), name + "_pc_feat": pc_feat.numpy(), name + "_pc": pc.numpy(),
                    name + "_add": add.numpy(), name + "_all_pcs": all_pcs.numpy()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_voxel_pe.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
