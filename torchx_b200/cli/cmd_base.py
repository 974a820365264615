import abc
import argparse


class SubCommand(abc.ABC):
    @abc.abstractmethod
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        raise NotImplementedError

    @abc.abstractmethod
    def run(self, args: argparse.Namespace) -> None:
        raise NotImplementedError
