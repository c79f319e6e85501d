"""Configuration handling: the reference's YAML files (config/network.yml, config/demo.yml) are
accepted unchanged.  The reference calls ``yaml.load(open(f))`` without a Loader (infer.py:280,
demo2_infer.py:58,65), which raises on PyYAML >= 6; ``yaml.safe_load`` reads the same files."""
import yaml


def load_config(path):
  with open(path) as f:
    return yaml.safe_load(f)


def cue_flags(config):
  """The five input-cue flags with the reference's defaults (infer.py:36-59)."""
  return {
      'use_depth': config.get('use_depth', True),
      'use_normals': config.get('use_normals', True),
      'use_class_probabilities': config.get('use_class_probabilities', False),
      'use_class_probabilities_pca': config.get('use_class_probabilities_pca', False),
      'use_intensity': config.get('use_intensity', False),
  }


SUPPORTED_LEGS = ('360OutputkLegs', '360OutputkLegsFixed')       # generateNet.py:119,222
SUPPORTED_OVERLAP_HEADS = ('DeltaLayerConv1NetworkHead',)        # generateNet.py:64
SUPPORTED_ORIENTATION_HEADS = ('CorrelationHead',)               # generateNet.py:327


def check_model(model_cfg):
  """The reference resolves legsType / overlap_head / orientation_head with getattr on
  generateNet (infer.py:91-93); an unknown name raises AttributeError there.  Same here."""
  for key, allowed in (('legsType', SUPPORTED_LEGS), ('overlap_head', SUPPORTED_OVERLAP_HEADS),
                       ('orientation_head', SUPPORTED_ORIENTATION_HEADS)):
    name = model_cfg[key]
    if name not in allowed:
      raise AttributeError("module 'generateNet' has no attribute 'generate%s'" % name)
