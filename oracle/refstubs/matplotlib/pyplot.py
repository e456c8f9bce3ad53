"""Empty pyplot stand-in; plotting is out of scope."""
