"""Fake-module injection for importing the plugin outside LiveTalking (the technique of the reference's own
tests/test_asr_server.py:57-72): minimal stand-ins for avatars.base_avatar / registry / utils.* so that
livetalking_b200.plugin.* can be imported on a box without the reference tree (and without av/aiortc/resampy...)."""
import logging
import sys
import types
from dataclasses import dataclass, field
from queue import Queue

import numpy as np


def install():
    if "avatars.base_avatar" in sys.modules and getattr(sys.modules["avatars.base_avatar"], "_LTB_STUB", False):
        return
    avatars = types.ModuleType("avatars")
    avatars.__path__ = []
    base = types.ModuleType("avatars.base_avatar")
    base._LTB_STUB = True

    @dataclass
    class AudioFrameData:                       # avatars/base_avatar.py:56-61
        data: np.ndarray
        type: int = 0
        userdata: dict = field(default_factory=dict)

    class BaseAvatar:                           # the attributes the plugin relies on (avatars/base_avatar.py:63-86)
        def __init__(self, opt):
            self.opt = opt
            self.sample_rate = 16000
            self.chunk = self.sample_rate // (opt.fps * 2)
            self.sessionid = getattr(opt, "sessionid", 0)
            self.custom_audiotype = 0
            self.custom_index = {}
            self.batch_size = opt.batch_size
            self.res_frame_queue = Queue(self.batch_size * 2)

        def get_avatar_length(self):
            return len(self.frame_list_cycle) if hasattr(self, "frame_list_cycle") else 1

    base.AudioFrameData = AudioFrameData
    base.BaseAvatar = BaseAvatar
    registry = types.ModuleType("registry")
    registry._REG = {}

    def register(category, name):
        def deco(cls):
            registry._REG.setdefault(category, {})[name] = cls
            return cls
        return deco

    def create(category, name, **kw):
        return registry._REG[category][name](**kw)

    registry.register, registry.create = register, create
    utils = types.ModuleType("utils")
    utils.__path__ = []
    image = types.ModuleType("utils.image")

    def mirror_index(size, index):              # utils/image.py:26-32
        turn, res = index // size, index % size
        return res if turn % 2 == 0 else size - res - 1

    def read_imgs(paths):
        import cv2
        return [cv2.imread(p) for p in paths]

    image.mirror_index, image.read_imgs = mirror_index, read_imgs
    logger_mod = types.ModuleType("utils.logger")
    logger_mod.logger = logging.getLogger("livetalking-stub")
    for name, mod in (("avatars", avatars), ("avatars.base_avatar", base), ("registry", registry), ("utils", utils),
                      ("utils.image", image), ("utils.logger", logger_mod)):
        sys.modules[name] = mod


class Opt:
    def __init__(self, batch_size=4, fps=25, l=10, r=10, **kw):
        self.batch_size, self.fps, self.l, self.r = batch_size, fps, l, r
        self.sessionid = 0
        for k, v in kw.items():
            setattr(self, k, v)
