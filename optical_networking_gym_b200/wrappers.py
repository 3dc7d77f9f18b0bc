"""`QRMSAEnvWrapper` with the reference's surface (wrappers/qrmsa_gym.py:24-87): forwards reset/step to the
B200 `QRMSAEnv`, caches `info["mask"]` and serves it through `action_masks()` (what sb3-contrib's MaskablePPO
calls), plus the helper pass-throughs the reference wrapper exposes."""
from __future__ import annotations

from .env import QRMSAEnv


class QRMSAEnvWrapper:
    metadata = {"render_modes": ["human"]}

    def __init__(self, *args, bands=None, **kwargs):
        if bands is not None:
            kwargs["bands"] = bands
        self.env = QRMSAEnv(*args, **kwargs)
        self.action_space = self.env.action_space
        self.observation_space = self.env.observation_space
        self.num_spectrum_resources = kwargs.get("num_spectrum_resources", 320)
        self.bit_rates = kwargs.get("bit_rates", (10, 40, 100))
        self.channel_width = kwargs.get("channel_width", 12.5)
        self.seed_value = kwargs.get("seed", 10)
        self._last_mask = None

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        if "mask" in info:
            self._last_mask = info["mask"]
        return obs, info

    def step(self, action):
        obs, reward, done, truncated, info = self.env.step(int(action))
        if "mask" in info:
            self._last_mask = info["mask"]
        return obs, reward, done, truncated, info

    def action_masks(self):
        return self._last_mask

    def render(self, mode="human"):
        pass

    def close(self):
        self.env.close()

    def get_available_slots(self, route):
        return self.env.get_available_slots(route)

    def get_number_slots(self, service, modulation):
        return self.env.get_number_slots(service, modulation)

    @property
    def unwrapped(self):
        return self.env
