"""Mirror of the reference's option handling (/root/reference/R/harmony_option.R).

``harmony_options()`` follows R/harmony_option.R:33-55 (same names with ``.`` -> ``_``, same
defaults); ``check_legacy_args`` follows :67-132 (legacy / unknown arguments are hard errors).
"""

LEGACY_ARGS = ("do_pca", "npcs", "tau", "block.size", "block_size", "max.iter.harmony", "max_iter_harmony",
               "max.iter.cluster", "max_iter_cluster", "epsilon.cluster", "epsilon_cluster", "epsilon.harmony",
               "epsilon_harmony")


class HarmonyOptions(dict):
    """The ``harmony_options`` S3 class of the reference (a named list)."""

    __getattr__ = dict.__getitem__


def validate_block_size(block_size):
    # R/harmony_option.R:58-63
    if block_size <= 0 or block_size > 1:
        raise ValueError("Error: block.size should be set between 0 and 1 (0 < block.size <= 1)")
    return block_size


def harmony_options(alpha=0.2, tau=0, block_size=0.05, max_iter_cluster=4, epsilon_cluster=1e-3,
                    epsilon_harmony=1e-2, batch_prop_cutoff=1e-5):
    # R/harmony_option.R:33-55
    block_size = validate_block_size(block_size)
    return HarmonyOptions(alpha=alpha, tau=tau, block_size=block_size, max_iter_cluster=max_iter_cluster,
                          epsilon_cluster=epsilon_cluster, epsilon_harmony=epsilon_harmony,
                          batch_prop_cutoff=batch_prop_cutoff)


def check_legacy_args(**kwargs):
    # R/harmony_option.R:67-81
    for arg in kwargs:
        if arg in LEGACY_ARGS:
            raise TypeError(
                f"Error: The parameter {arg} has been dropped from the RunHarmony API. Advanced users can set "
                f"its value through the .options parameter and harmony_options().")
    if kwargs:
        raise TypeError(
            f"Argument {', '.join(kwargs)} is unhandled. Please refer to the documentation for the valid "
            f"harmony options!")
