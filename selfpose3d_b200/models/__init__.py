"""Host-side mirror of the reference's ``lib/models`` package (same module and class names,
constructor/forward signatures and state-dict keys), backed by the sm_100a kernels of libsp3d."""
from . import pose_resnet  # noqa: F401
from . import v2v_net  # noqa: F401
from . import project_layer  # noqa: F401
from . import cuboid_proposal_net  # noqa: F401
from . import cuboid_proposal_net_soft  # noqa: F401
from . import pose_regression_net  # noqa: F401
from . import multi_person_posenet  # noqa: F401
from . import multi_person_posenet_ssv  # noqa: F401
