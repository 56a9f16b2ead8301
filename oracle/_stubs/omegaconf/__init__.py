"""Six-line stand-in for omegaconf (absent offline) so that /root/reference/src/global_cfg.py
imports; only DictConfig attribute access is needed by EpipolarTransformer.__init__."""


class DictConfig(dict):
    def __getattr__(self, k):
        v = self[k]
        return DictConfig(v) if isinstance(v, dict) else v


class OmegaConf:
    pass
