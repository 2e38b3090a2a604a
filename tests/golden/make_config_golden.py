"""Flattened key -> value tables of the reference's four benchmark yamls (configs/single-prompt_benchmark/asd_sd_nerf.yaml,
asd_mv_nerf.yaml, configs/multi-prompt_benchmark/asd_sd_hyper_iNGP_50k.yaml, asd_mv_triplane_transformer_10k.yaml), so that
tests/test_config_cpu.py can check that tests/configs/*.yaml still carry exactly the reference's keys and values
(interpolations are kept as their `${...}` strings). Run in the build container only: python tests/golden/make_config_golden.py"""
import json
import os

import yaml

REF = "/root/reference/configs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_golden.json")
PAIRS = {"asd_sd_nerf.yaml": "single-prompt_benchmark/asd_sd_nerf.yaml", "asd_mv_nerf.yaml": "single-prompt_benchmark/asd_mv_nerf.yaml",
         "asd_sd_hyper_iNGP.yaml": "multi-prompt_benchmark/asd_sd_hyper_iNGP_50k.yaml",
         "asd_mv_triplane_transformer.yaml": "multi-prompt_benchmark/asd_mv_triplane_transformer_10k.yaml"}


def flat(d, pre=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(flat(v, pre + str(k) + "."))
        else:
            out[pre + str(k)] = v
    return out


if __name__ == "__main__":
    json.dump({mine: flat(yaml.safe_load(open(os.path.join(REF, ref)))) for mine, ref in PAIRS.items()}, open(OUT, "w"),
              indent=0, sort_keys=True)
    print("wrote", OUT)
