"""pips-ipmpp_b200: B200-native KKT factor/solve hot path of PIPS-IPM++.

The directory name contains a hyphen, so import it through `__graft_entry__.load_package()` (module name
`pips_ipmpp_b200`)."""
from . import problem  # noqa: F401
from . import capi  # noqa: F401
from . import linsys  # noqa: F401
from . import regularization  # noqa: F401
