"""serenity_b200 - B200-native XC / embedding-potential build behind Serenity's FuncPotential interface."""
__version__ = "0.1.0"
