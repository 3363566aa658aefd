"""pyxsim_b200 -- B200-native (sm_100a) photon generation and projection behind pyXSIM's
``make_photons`` / ``project_photons`` / ``SourceModel`` API.

Public names follow ``pyxsim/__init__.py:11-31`` for the parts that belong to the hot path.
"""
from .data import ArrayDataSource, Cosmology
from .event_list import EventList, merge_files
from .internal_absorption import (
    column_density_at_cells,
    make_column_density_map,
    read_column_file,
    write_column_file,
)
from .photon_list import (
    DeviceList,
    drop_list,
    load_list,
    make_events,
    make_photons,
    project_photons,
    project_photons_allsky,
    set_async_writes,
    wait_for_writes,
)
from .source_models import (
    CIESourceModel,
    IGMSourceModel,
    LineSourceModel,
    NEISourceModel,
    PionSourceModel,
    PowerLawSourceModel,
    SourceModel,
    ThermalSourceModel,
)
from .spectral_models import (
    AbsorptionModel,
    PionSpectralModel,
    TableCIEModel,
    TBabsModel,
    ThermalSpectralModel,
    WabsModel,
    absorb_models,
)

__version__ = "0.1.0"
