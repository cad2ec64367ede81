"""Minimal stand-in for ``yacs.config.CfgNode`` (yacs is not installed and there is no network).

TEST INFRASTRUCTURE ONLY: lets the unmodified reference under /root/reference be imported in this
container so oracle/ can be validated against it and golden fixtures generated.  Never imported by
the product package.
"""
import copy
import yaml
from ast import literal_eval


class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v
        self.__dict__['_frozen'] = False

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if isinstance(value, dict) and not isinstance(value, CfgNode):
            value = CfgNode(value)
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            out[k] = copy.deepcopy(v, memo)
        return out

    def defrost(self):
        pass

    def freeze(self):
        pass

    def is_frozen(self):
        return False

    @classmethod
    def load_cfg(cls, f):
        if hasattr(f, 'read'):
            f = f.read()
        return cls(yaml.safe_load(f))

    def merge_from_other_cfg(self, other):
        for k, v in other.items():
            if k not in self:
                raise KeyError(f'Non-existent config key: {k}')
            if isinstance(v, dict) and isinstance(self[k], dict):
                self[k].merge_from_other_cfg(v)
            else:
                self[k] = copy.deepcopy(v)

    def merge_from_file(self, fname):
        with open(fname) as f:
            self.merge_from_other_cfg(CfgNode.load_cfg(f))

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split('.')
            for p in parts[:-1]:
                node = node[p]
            if parts[-1] not in node:
                raise KeyError(f'Non-existent config key: {k}')
            if isinstance(v, str):
                try:
                    v = literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
            old = node[parts[-1]]
            if isinstance(old, float) and isinstance(v, int):
                v = float(v)
            node[parts[-1]] = v
