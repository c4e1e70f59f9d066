"""``cocos_b200.scientific``: the Gaussian KDE that follows the Heston simulation in the reference's workflow."""
from .kde import evaluate_gaussian_kde_in_batches, gaussian_kde   # noqa: F401
