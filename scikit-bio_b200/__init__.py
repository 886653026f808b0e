"""B200-native all-pairs beta-diversity engine (drop-in for skbio.diversity.beta_diversity).

Import as ``skbio_b200`` (see ``skbio_b200/__init__.py``).
"""
from ._tree import (SimpleTreeNode, FlatTree, read_newick, read_newick_flat, flatten_tree,  # noqa: F401
                    validate_taxa_and_tree, DuplicateNodeError, MissingNodeError)
from ._dm import DistanceMatrix, SimpleDistanceMatrix  # noqa: F401
from ._capi import EngineError, available, release_cache, set_cache_limit  # noqa: F401
from ._driver import (alpha_diversity, beta_diversity, block_beta_diversity,  # noqa: F401
                      partial_beta_diversity, install, last_call_profile)

from ._stats import center_distance_matrix, permanova  # noqa: F401
from ._io import write_binary_dm, write_lsmat  # noqa: F401

__version__ = "0.1.0"
