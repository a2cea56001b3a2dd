__version__ = "2.1.5+b200.r1"
