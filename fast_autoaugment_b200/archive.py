"""Policy archive with the reference's function surface.

Mirrors ``FastAutoAugment/archive.py``: every function returns
``list[list[(op_name, prob, level)]]`` - the format ``Augmentation(policy)``
consumes (reference ``data.py:254``) - so ``train.py`` / ``search.py`` style
callers need no change.  The searched numbers themselves are data, stored under
``policies/*.json`` (written by ``tools/export_policies.py``); levels are the
post-``autoaug2arsaug`` (reference ``archive.py:59-87``) values, i.e. exactly
what the reference functions return.
"""
from __future__ import annotations

import functools
import json
import os

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "policies")

# op order of the reference's search space: augment_list(False),
# reference augmentations.py:157-173 - indices used by ``policy_decoder``.
SEARCH_OPS = ("ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate", "AutoContrast",
              "Invert", "Equalize", "Solarize", "Posterize", "Contrast", "Color",
              "Brightness", "Sharpness", "Cutout")


@functools.lru_cache(maxsize=None)
def _load(name):
    with open(os.path.join(_DIR, name + ".json")) as f:
        d = json.load(f)
    k = d["n_op"]
    rows = d["table"]
    return tuple(tuple((d["ops"][int(rows[s * k + j][0])], rows[s * k + j][1], rows[s * k + j][2])
                       for j in range(k)) for s in range(d["n_sub"]))


def _as_lists(name):
    return [[list(op) if name.startswith("fa_") else tuple(op) for op in sub] for sub in _load(name)]


def fa_reduced_cifar10():        # reference archive.py:281
    return _as_lists("fa_reduced_cifar10")


def fa_resnet50_rimagenet():     # reference archive.py:286
    return _as_lists("fa_resnet50_rimagenet")


def fa_reduced_svhn():           # reference archive.py:291
    return _as_lists("fa_reduced_svhn")


def arsaug_policy():             # reference archive.py:11
    return _as_lists("arsaug_policy")


def autoaug_paper_cifar10():     # reference archive.py:90-119
    return _as_lists("autoaug_paper_cifar10")


def autoaug_policy():            # reference archive.py:122-242
    return _as_lists("autoaug_policy")


def remove_deplicates(policies):
    """Keep the first sub-policy for every distinct op-name sequence
    (reference archive.py:264-278; the reference's spelling is kept)."""
    seen, kept = set(), []
    for sub in policies:
        key = "_".join(op[0] for op in sub)
        if key not in seen:
            seen.add(key)
            kept.append(sub)
    return kept


def policy_decoder(augment, num_policy, num_op):
    """hyperopt sample dict -> policy list (reference archive.py:296-307)."""
    return [[(SEARCH_OPS[augment["policy_%d_%d" % (i, j)]],
              augment["prob_%d_%d" % (i, j)],
              augment["level_%d_%d" % (i, j)]) for j in range(num_op)]
            for i in range(num_policy)]


# the names ``get_dataloaders`` accepts for conf['aug'] (reference data.py:91-105)
BY_CONF_NAME = {
    "fa_reduced_cifar10": fa_reduced_cifar10,
    "fa_reduced_imagenet": fa_resnet50_rimagenet,
    "fa_reduced_svhn": fa_reduced_svhn,
    "arsaug": arsaug_policy,
    "autoaug_cifar10": autoaug_paper_cifar10,
    "autoaug_extend": autoaug_policy,
}
