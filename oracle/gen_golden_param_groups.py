"""tests/golden/fai_detr_l_param_groups.json: the optimiser parameter groups the UNMODIFIED reference builds for
fai-detr-l-obj365 (focoos/trainer/solver/build.py:40-101 with the trainer defaults), as [name, lr, weight_decay].

    python -m oracle.gen_golden_param_groups
"""
import json
import os

from oracle import ref_import


def main():
    fm = ref_import.get_reference_model("fai-detr-l-obj365")
    from focoos.trainer.solver.build import get_optimizer_params

    model = fm.model
    names = {id(p): n for n, p in model.named_parameters()}
    groups = get_optimizer_params(model, base_lr=1e-4, weight_decay=1e-4, weight_decay_norm=0.0, weight_decay_embed=0.0, backbone_multiplier=0.1,
                                  decoder_multiplier=1.0, head_multiplier=1.0)
    rows = [[names[id(g["params"][0])], g["lr"], g["weight_decay"], g["params"][0].numel()] for g in groups]
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fai_detr_l_param_groups.json")
    json.dump({"base_lr": 1e-4, "weight_decay": 1e-4, "groups": rows, "total_elements": sum(r[3] for r in rows)}, open(out, "w"))
    print(len(rows), "groups,", sum(r[3] for r in rows), "elements ->", out)


if __name__ == "__main__":
    main()
