class OmegaConf:
    @staticmethod
    def resolve(cfg):
        return cfg
