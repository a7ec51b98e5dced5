"""The engine wrapper lives in the package (assistedmanipulation_b200/engine.py); the tests import it through here."""
from assistedmanipulation_b200.engine import Engine, lib, ptr  # noqa: F401
