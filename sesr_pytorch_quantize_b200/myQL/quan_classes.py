"""Value objects that describe "insert this function around modules of that type"
(mirror of the reference's myQL/quan_classes.py:9-38: same class names, constructor arguments, methods)."""
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Optional, Tuple, Type


@dataclass
class FunctionPackage:
    """A callable plus the keyword arguments every inserted call gets."""
    function: Callable
    parameter_dict: Optional[Dict[str, Any]] = None

    def __init__(self, the_function: Callable, parameter_dict: Optional[Dict[str, Any]] = None):
        self.function = the_function
        self.parameter_dict = parameter_dict


@dataclass
class NodeInsertConfig:
    should_insert: bool
    function_package: Optional[FunctionPackage] = None


class NodeInsertMappingElement:
    """(module type, FunctionPackage) pair."""

    def __init__(self, insert_type, the_function: FunctionPackage):
        self.insert_mapping_config = (insert_type, the_function)

    def get_config(self) -> Tuple[Type, FunctionPackage]:
        return self.insert_mapping_config


class NodeInsertMapping:
    """module type -> FunctionPackage."""

    def __init__(self):
        self.insert_mapping: Dict[Type, FunctionPackage] = {}

    def add_config(self, insert_mapping_config: NodeInsertMappingElement):
        kind, package = insert_mapping_config.get_config()
        self.insert_mapping[kind] = package

    def get_mapping(self):
        return self.insert_mapping
