"""Mirror of the reference's ``model`` package for the hot path: put this package's parent
directory first on ``sys.path`` and ``from model.ifnet import IFNet`` / ``from model.projection
import project`` resolve to the B200-native implementations with unchanged signatures."""
