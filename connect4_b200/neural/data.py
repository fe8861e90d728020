"""Generation sink with the reference's on-disk schema (oinkoink/neural/pytorch/data.py:13-105):
`data.pth` = {'boards': f32 [2P,3,6,7], 'values': f32 [2P], 'priors': f32 [2P,7]}, originals first, mirrors second.
The flip augmentation and plane expansion run on the device (c4_records_augment_pack)."""
import numpy as np
import torch

from ..engine import RECORD_DTYPE, augment_pack
from .storage import GameStorage


class Connect4Dataset(torch.utils.data.Dataset):
    def __init__(self, boards, values, priors):
        self.boards = boards
        self.values = values
        self.priors = priors

    def save(self, filename):
        torch.save({'boards': self.boards, 'values': self.values, 'priors': self.priors}, filename)

    @classmethod
    def load(cls, filename):
        data = torch.load(filename)
        return cls(data['boards'], data['values'], data['priors'])

    def __len__(self):
        return len(self.boards)

    def __getitem__(self, idx):
        if self.priors is None:
            return self.boards[idx], self.values[idx]
        return self.boards[idx], self.values[idx], self.priors[idx]


def records_from_lists(boards, values, priors):
    rec = np.zeros(len(boards), dtype=RECORD_DTYPE)
    rec["c0"] = [int(b.color[0]) for b in boards]
    rec["c1"] = [int(b.color[1]) for b in boards]
    rec["result_value"] = np.asarray(values, np.float32)
    rec["policy"] = np.asarray(priors, np.float32).reshape(len(boards), 7)
    return rec


def native_to_pytorch(boards, values, priors=None, to_move_channel=True, add_fliplr=False):
    """data.py:78-105 on the device. Unlike the reference it does not extend the caller's lists in place."""
    assert len(boards) == len(values)
    pri = priors if priors is not None else np.zeros((len(boards), 7))
    rec = torch.as_tensor(records_from_lists(boards, values, pri).view(np.uint8).reshape(-1, 64)).cuda()
    b, v, p = augment_pack(rec)
    if not add_fliplr:
        n = len(boards)
        b, v, p = b[:n], v[:n], p[:n]
    if not to_move_channel:
        b = b[:, 1:]
    return b.cpu(), v.cpu(), (p.cpu() if priors is not None else None)


class TrainingDataStorage(GameStorage):
    def td_file_name(self, folder_path, gen):
        return "{}/{}/data.pth".format(folder_path, gen)

    def save(self, games, folder_path):
        super().save(games, folder_path)
        data = sum(g.data for g in games)
        board_t, value_t, prior_t = native_to_pytorch(data.boards, data.values, data.priors, add_fliplr=True)
        Connect4Dataset(board_t, value_t, prior_t).save(folder_path + '/data.pth')

    def save_records(self, records_device, folder_path):
        """straight from the device records of a generation (no per-board Python objects)"""
        b, v, p = augment_pack(records_device)
        Connect4Dataset(b.cpu(), v.cpu(), p.cpu()).save(folder_path + '/data.pth')
