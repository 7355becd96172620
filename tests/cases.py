"""The synthetic input decks live in ``snac_b200/decks.py`` (they are also what ``bench.py`` and
``smoke()`` run); the tests keep importing them under this name."""
from snac_b200.decks import *   # noqa: F401,F403
from snac_b200.decks import write_case, make, random_fields, channel_case, cavity_case, tgv_case, duct_blocks_case, developing_channel_case, lshape_blocks_case   # noqa: F401
