"""Minimal stand-in for the three TensorFlow calls the reference's ``train.py`` / ``seeding.py`` make at import and seed
time (``/root/reference/sac_eo/train.py:17,29-31``, ``common/seeding.py:3,13``).  Put this directory on PYTHONPATH
ONLY to drive the reference's unmodified ``train.py`` against the ``sac_eo`` package of this repository
(INTEGRATION.md); every numerical call of the reference's own TensorFlow classes is replaced by libsaceo, so nothing
else of TensorFlow is needed - and nothing else is provided: any other attribute access fails loudly."""


class _Experimental:
    @staticmethod
    def list_physical_devices(kind=None):
        return []

    @staticmethod
    def set_memory_growth(device, enable):
        return None


class config:
    experimental = _Experimental


class random:
    @staticmethod
    def set_seed(seed):
        return None


def __getattr__(name):
    raise AttributeError("shims/tensorflow only provides config.experimental.{list_physical_devices,set_memory_growth} "
                         "and random.set_seed; '%s' is not on the libsaceo path" % name)
