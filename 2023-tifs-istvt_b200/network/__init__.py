"""Host-side mirror of the reference's `network` package for the ISTVT path (same module/attribute names,
so `state_dict` keys, `model_selection(...)` and `XceptionVidTr()` call sites keep working)."""
