"""SM partitions for callers that keep several batches in flight (C ABI: pn2_sm_partition_*).

    part = SmPartition(fps_sms=80)            # sampling chains on >= 80 SMs, everything else on the rest
    lane = part.stream(SmPartition.MAIN)      # a caller stream: ball query, fused MLP kernels, FP
    net.sm_partition = part                   # the backbone takes its sampling side streams from the other group

PyTorch only wraps the stream handles (torch.cuda.ExternalStream); the partition is a pair of CUDA green contexts.
"""
import ctypes
import os

import torch

from ._lib import check, lib


def lane_stream(device):
    """A caller ("lane") stream for one batch in flight: the short kernels of a step (ball query, fused MLPs, FP).
    PN2_LANE_PRIORITY (tuning, default 0).  Measured on the 8-lane bench step: raising the lanes above the sampling
    streams (-1) costs 8 % (10.8 k against 11.8 k scenes/s) -- the sampling chains are the long pole of every lane, and
    delaying their launches behind other lanes' short kernels lengthens all of them."""
    return torch.cuda.Stream(device=device, priority=int(os.environ.get("PN2_LANE_PRIORITY", "0")))


def sampling_stream(device):
    """The side stream a lane's FPS pyramid runs on (PN2_SAMPLING_PRIORITY, default 0; negative = above the lanes)."""
    return torch.cuda.Stream(device=device, priority=int(os.environ.get("PN2_SAMPLING_PRIORITY", "0")))


class SmPartition:
    FPS, MAIN = 0, 1

    def __init__(self, fps_sms, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()
            check(lib.pn2_sm_partition_create(int(fps_sms), ctypes.byref(self._handle)), "sm_partition_create")
        self.sms = (lib.pn2_sm_partition_sms(self._handle, 0), lib.pn2_sm_partition_sms(self._handle, 1))
        self._streams = []

    def stream(self, which):
        """A new non-blocking stream whose kernels run on the SMs of group ``which`` only."""
        s = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.pn2_sm_partition_stream_create(self._handle, int(which), ctypes.byref(s)), "sm_partition_stream_create")
        st = torch.cuda.ExternalStream(s.value, device=self.device)
        self._streams.append(st)
        return st
