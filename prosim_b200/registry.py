"""Class registry with the reference's decorator / getter names (prosim/core/registry.py:25-136).

The reference asserts ``LightningModule`` for models; pytorch_lightning is not a dependency of this
path, so models are checked against ``nn.Module`` (a LightningModule is one).  To plug ProSimB200 into
the reference's own registry instead, see INTEGRATION.md.
"""
import collections

from torch import nn


class Registry:
    mapping = collections.defaultdict(dict)

    @classmethod
    def _register_impl(cls, _type, to_register, name, assert_type=None):
        def wrap(to_register):
            if assert_type is not None:
                assert issubclass(to_register, assert_type), f'{to_register} must be a subclass of {assert_type}'
            cls.mapping[_type][to_register.__name__ if name is None else name] = to_register
            return to_register
        return wrap if to_register is None else wrap(to_register)

    @classmethod
    def register_model(cls, to_register=None, *, name=None):
        return cls._register_impl('model', to_register, name, assert_type=nn.Module)

    @classmethod
    def register_scene_encoder(cls, to_register=None, *, name=None):
        return cls._register_impl('scene_encoder', to_register, name, assert_type=nn.Module)

    @classmethod
    def register_prompt_encoder(cls, to_register=None, *, name=None):
        return cls._register_impl('prompt_encoder', to_register, name, assert_type=nn.Module)

    @classmethod
    def register_decoder(cls, to_register=None, *, name=None):
        return cls._register_impl('decoder', to_register, name, assert_type=nn.Module)

    @classmethod
    def register_policy(cls, to_register=None, *, name=None):
        return cls._register_impl('policy', to_register, name, assert_type=nn.Module)

    @classmethod
    def _get_impl(cls, _type, name):
        return cls.mapping[_type].get(name, None)

    @classmethod
    def get_model(cls, name):
        return cls._get_impl('model', name)

    @classmethod
    def get_scene_encoder(cls, name):
        return cls._get_impl('scene_encoder', name)

    @classmethod
    def get_prompt_encoder(cls, name):
        return cls._get_impl('prompt_encoder', name)

    @classmethod
    def get_decoder(cls, name):
        return cls._get_impl('decoder', name)

    @classmethod
    def get_policy(cls, name):
        return cls._get_impl('policy', name)


registry = Registry()
