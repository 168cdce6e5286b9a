"""Roster loader — mirror of src/ingest.py:611-663 (ElectorDataIngester) plus the conversion of
the three columns the engine reads (src/model.py:199-213) into device-uploadable arrays."""
from __future__ import annotations

import logging
from pathlib import Path
from typing import Tuple

import numpy as np
import pandas as pd


class ElectorDataIngester:
    """Same constructor / method / return contract as the reference: a DataFrame indexed by
    'elector_id' (as str), or an EMPTY DataFrame on any failure (which is logged)."""

    def __init__(self, elector_data_file: str):
        self.elector_data_file_path = Path(elector_data_file)
        self.log = logging.getLogger(__name__ + ".ElectorDataIngester")

    def load_and_prepare_data(self) -> pd.DataFrame:
        path = self.elector_data_file_path
        self.log.info(f"Attempting to load elector data from: {path}")
        if not path.exists():
            self.log.error(f"Elector data file not found: {path}")
            return pd.DataFrame()
        try:
            df = pd.read_csv(path)
        except pd.errors.EmptyDataError:
            self.log.error(f"Elector data file is empty or malformed (EmptyDataError): {path}")
            return pd.DataFrame()
        except Exception as exc:  # same catch-all as the reference
            self.log.error(f"An unexpected error occurred while loading elector data from {path}: {exc}", exc_info=True)
            return pd.DataFrame()
        if df.empty:
            self.log.warning(f"Elector data file is empty: {path}")
            return pd.DataFrame()
        if "elector_id" not in df.columns:
            if "gc_id" not in df.columns:
                self.log.error("'elector_id' (and 'gc_id' fallback) column not found in elector data.")
                return pd.DataFrame()
            self.log.warning("'elector_id' column not found, using 'gc_id' as elector_id.")
            df["elector_id"] = df["gc_id"]
        df["elector_id"] = df["elector_id"].astype(str)
        df = df.set_index("elector_id")
        self.log.info(f"Successfully loaded and prepared {len(df)} records from {path}.")
        return df


def roster_arrays(elector_data: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(ideology f64, region codes int32, papabile bool) in roster order.  Region labels are
    factorised: the engine only ever tests two electors' regions for equality (model.py:322).
    A missing 'is_papabile' column means everybody is papabile (model.py:211-215)."""
    n = len(elector_data)
    x = np.asarray(elector_data["ideology_score"].values, dtype=np.float64)
    labels = list(elector_data["region"].values)
    codes: dict = {}
    g = np.empty(n, np.int32)
    for i, lab in enumerate(labels):
        key = ("nan", i) if (isinstance(lab, float) and lab != lab) else lab    # NaN equals nothing, itself included
        g[i] = codes.setdefault(key, len(codes))
    if "is_papabile" in elector_data:
        pap = np.asarray(elector_data["is_papabile"].astype(bool).values, dtype=bool)
    else:
        pap = np.ones(n, bool)
    return x, g, pap
