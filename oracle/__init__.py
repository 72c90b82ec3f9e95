"""Test infrastructure only (CPU oracle). Never imported by vits_b200/."""
