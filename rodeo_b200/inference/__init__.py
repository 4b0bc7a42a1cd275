from .basic import basic
from .fenrir import fenrir
from .dalton import dalton
