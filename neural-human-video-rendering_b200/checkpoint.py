"""Checkpoint naming of the reference: ``<checkpoints_dir>/<name>/<epoch>_net_<label>.pth`` state_dicts
[REF start.sh:7-8,28 --name/--checkpoints_dir/--which_epoch; pretrain_start.sh:29-30
--load_pretrain_TransG/--which_epoch_TransG; file naming per pix2pixHD BaseModel.save_network]."""
from __future__ import annotations

import os

import torch

LABELS = {"G": "netG", "TransG": "netTransG", "BG": "netBG"}


def net_path(save_dir: str, epoch, label: str) -> str:
    return os.path.join(save_dir, "%s_net_%s.pth" % (epoch, label))


def save_pipeline(pipe, save_dir: str, epoch) -> None:
    os.makedirs(save_dir, exist_ok=True)
    for label, attr in LABELS.items():
        torch.save(getattr(pipe, attr).state_dict(), net_path(save_dir, epoch, label))
    torch.save({"atlas": pipe.atlas.detach().cpu(), "bg": pipe.bg.detach().cpu()}, net_path(save_dir, epoch, "Tex"))


def load_pipeline(pipe, save_dir: str, epoch, only=None) -> bool:
    """Loads whatever exists; returns True if at least one file was found."""
    found = False
    for label, attr in LABELS.items():
        if only and label not in only:
            continue
        p = net_path(save_dir, epoch, label)
        if os.path.isfile(p):
            getattr(pipe, attr).load_state_dict(torch.load(p, map_location="cpu"))
            found = True
    p = net_path(save_dir, epoch, "Tex")
    if (not only or "Tex" in only) and os.path.isfile(p):
        d = torch.load(p, map_location="cpu")
        with torch.no_grad():
            pipe.atlas.copy_(d["atlas"])
            pipe.bg.copy_(d["bg"])
        found = True
    return found
