"""Import stub: the reference's target modules import matplotlib at module level for their plotting metrics; nothing on the
hot path draws."""
