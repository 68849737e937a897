"""matplotlib is imported by updes/utils.py for plotting helpers only; nothing on the hot path draws."""
