"""Host-side mirror of the batch containers the rollout path consumes.

Same names, attribute access and ``__to__`` convention as the reference's
``InputMaskData / BatchDataDict / BatchPrompt`` (prosim/dataset/format_utils.py:30-145)
and ``BatchCondition`` (prosim/dataset/condition_utils.py:103-124), so a batch built
for the reference model can be handed to ``ProSimB200.forward`` unchanged and the
other way round.  Layouts:

  init_obs / fut_obs[t]   input [B,A,11,24] f32, mask [B,A,11,24] bool,
                          position [B,A,2], heading [B,A], agent_ids list[list[str]]
  init_map                input [B,M,19,11], mask [B,M,19], position [B,M,1,2], heading [B,M,1]
  prompt['motion_pred']   prompt [B,A,7], prompt_mask [B,A], position [B,A,2],
                          heading [B,A,1], agent_type [B,A] int64, agent_ids
  condition['goal']       input [B,C,3], mask [B,C], prompt_idx [B,C,1], prompt_mask [B,A]
"""
import torch

_IMD_KEYS = ('input', 'mask', 'position', 'heading', 'agent_ids')


class InputMaskData:
    def __init__(self, input, mask, position=None, heading=None, agent_ids=None):
        self.input = input
        self.mask = mask
        self.position = position
        self.heading = heading
        self.agent_ids = agent_ids
        # same guard as the reference: no NaN may sit under a True mask
        assert not bool(input[mask].isnan().any())

    @classmethod
    def from_dict(cls, data_dict):
        return cls(**data_dict)

    def __to__(self, device, non_blocking=False):
        self.input = self.input.to(device, non_blocking=non_blocking)
        self.mask = self.mask.to(device, non_blocking=non_blocking)
        if self.position is not None:
            self.position = self.position.to(device, non_blocking=non_blocking)
            self.heading = self.heading.to(device, non_blocking=non_blocking)
        return self

    def __setitem__(self, key, value):
        assert key in _IMD_KEYS
        setattr(self, key, value)

    def __getitem__(self, key):
        assert key in _IMD_KEYS
        return getattr(self, key)

    def keys(self):
        keys = ['input', 'mask']
        if self.position is not None:
            keys += ['position', 'heading']
        if self.agent_ids is not None:
            keys += ['agent_ids']
        return keys


class BatchDataDict:
    def __init__(self, input_dict):
        self.input = input_dict

    def __to__(self, device, non_blocking=False):
        for key, val in self.input.items():
            if isinstance(val, list):
                continue
            if isinstance(val, torch.Tensor):
                self.input[key] = val.to(device, non_blocking=non_blocking)
            else:
                self.input[key] = val.__to__(device, non_blocking=non_blocking)
        return self

    def __getitem__(self, key):
        return self.input[key]

    def keys(self):
        return self.input.keys()


class _DictOfDicts:
    _skip = (list, str)

    def __init__(self, data):
        self._data = data

    def __to__(self, device, non_blocking=False):
        for sub in self._data.values():
            for k, v in sub.items():
                if isinstance(v, self._skip):
                    continue
                sub[k] = v.to(device, non_blocking=non_blocking)
        return self

    def __getitem__(self, key):
        assert key in self._data.keys()
        return self._data[key]

    def __len__(self):
        return len(self._data)

    def keys(self):
        return self._data.keys()


class BatchPrompt(_DictOfDicts):
    @property
    def all_prompts(self):
        return self._data

    @all_prompts.setter
    def all_prompts(self, value):       # a plain attribute in the reference: callers assign to it
        self._data = value


class BatchCondition(_DictOfDicts):
    @property
    def all_cond(self):
        return self._data

    @all_cond.setter
    def all_cond(self, value):          # rollout/gpu_utils.py:175 replaces the whole condition set
        self._data = value


class SceneBatch:
    """The two attributes of trajdata's SceneBatch the rollout path touches."""

    def __init__(self, scene_ids, extras):
        self.scene_ids = scene_ids
        self.extras = extras

    def to(self, device, non_blocking=False):
        for key, val in self.extras.items():
            if isinstance(val, torch.Tensor):
                self.extras[key] = val.to(device, non_blocking=non_blocking)
            elif hasattr(val, '__to__'):
                self.extras[key] = val.__to__(device, non_blocking=non_blocking)
        return self
