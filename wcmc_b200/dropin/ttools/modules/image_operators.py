from sbmc.modules import crop_like  # noqa: F401
