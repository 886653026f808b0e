"""B200-native all-pairs beta-diversity engine (drop-in for skbio.diversity.beta_diversity).

Import as ``skbio_b200`` (see ``skbio_b200/__init__.py``).
"""
from ._tree import (SimpleTreeNode, read_newick, flatten_tree,  # noqa: F401
                    DuplicateNodeError, MissingNodeError)

__version__ = "0.1.0"
