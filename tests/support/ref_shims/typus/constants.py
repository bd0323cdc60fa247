import enum


class RankLevel(enum.Enum):
    """Taxonomic rank levels (the reference only uses .value, .name, hashing and ordering by value)."""

    L5 = 5
    L10 = 10
    L11 = 11
    L12 = 12
    L13 = 13
    L15 = 15
    L20 = 20
    L24 = 24
    L25 = 25
    L26 = 26
    L27 = 27
    L30 = 30
    L32 = 32
    L33 = 33
    L33_5 = 335
    L34 = 34
    L34_5 = 345
    L35 = 35
    L37 = 37
    L40 = 40
    L43 = 43
    L44 = 44
    L45 = 45
    L47 = 47
    L50 = 50
    L53 = 53
    L57 = 57
    L60 = 60
    L67 = 67
    L70 = 70
