"""Golden evaluation batches of the multi-prompt data modules from the REFERENCE's own classes
(custom/amortized/data/multiprompt.py: MultipromptRandomCameraDataset4Test :85-122, MultipromptRandomCameraDataset4FixPrompt
:125-164), extracted and executed unchanged the way make_data_golden.py does it.

Run in the build container only (it reads /root/reference):
    python tests/golden/make_eval_data_golden.py
`torch.manual_seed(s)` precedes each construction (the 4Test set draws its noise end points there). The FixPrompt set is
collated with torch's default_collate at batch size 1, as its DataLoader does (multiprompt.py:230-234).
Output: tests/golden/eval_data_golden.pt (camera scalars only; rays are generated on the device by the product).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_data_golden import namespace, pieces  # noqa: E402

OUT = os.path.join(HERE, "eval_data_golden.pt")
DROP = ("rays_o", "rays_d")


def keep(batch):
    return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items() if k not in DROP}


def main():
    ns = namespace()
    pieces("/root/reference/custom/amortized/data/multiprompt.py",
           ["MultipromptRandomCameraDataModuleConfig", "MultipromptRandomCameraDataset4Test",
            "MultipromptRandomCameraDataset4FixPrompt"], ns)
    from torch.utils.data._utils.collate import default_collate

    base = dict(eval_height=12, eval_width=20, n_val_views=4, n_test_views=6, eval_elevation_deg=15.0,
                eval_camera_distance=3.0, eval_fovy_deg=40.0, dim_gaussian=8)
    library = {"train": ["a", "b", "c"], "val": ["a red apple", "a wooden chair"], "test": ["a blue car"]}
    gold = {"library": library, "cases": []}
    for split, seed in (("val", 5), ("test", 6)):
        torch.manual_seed(seed)
        ds = ns["MultipromptRandomCameraDataset4Test"](ns["MultipromptRandomCameraDataModuleConfig"](**base), split, library)
        gold["cases"].append({"kind": "library", "split": split, "seed": seed, "config": base,
                              "batches": [keep(ds.collate(item)) for item in ds]})
    for extra in (dict(eval_prompt="a corgi"), dict(eval_prompt="a corgi", target_prompt="a cat"),
                  dict(eval_prompt="a corgi", eval_fix_camera=2), dict(eval_prompt="a corgi", eval_fix_camera=0)):
        cfg_kw = dict(base, **extra)
        ds = ns["MultipromptRandomCameraDataset4FixPrompt"](ns["MultipromptRandomCameraDataModuleConfig"](**cfg_kw), "test")
        gold["cases"].append({"kind": "fix_prompt", "split": "test", "config": cfg_kw,
                              "batches": [keep(default_collate([item])) for item in ds]})
    torch.save(gold, OUT)
    print("wrote", OUT, [(c["kind"], len(c["batches"])) for c in gold["cases"]])
    print({k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in gold["cases"][0]["batches"][0].items()})
    print({k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in gold["cases"][3]["batches"][1].items()})


if __name__ == "__main__":
    main()
