{-|
Runs the UNMODIFIED reference (qubitrot/qubism) on the golden cases of this repository and prints
its amplitudes -- the missing pin of oracle/ (README.md in this directory).  Compiled inside the
reference's stack project: it imports the reference's own modules, nothing of this repository.

  GATE id n U q th ph la | re im re im ...        onJust q (unitary th ph la) #> v
  GATE id n CU c t th ph la | ...                 controlled c (onJust t (unitary ..)) #> v
  GATE id n CX c t | ...                          cnot c t #> v
  GATE id n COLLAPSE q b | ...                    collapse q b v
  PROG name run file draws...                     parseOpenQASM + runStmt with forced draws
-}
{-# LANGUAGE DataKinds, KindSignatures, RankNTypes, ScopedTypeVariables, GeneralizedNewtypeDeriving #-}
module Main where

import GHC.TypeLits
import Data.Proxy
import Data.Finite
import Data.Complex
import Data.List (intercalate)
import qualified Data.Map.Strict as Map
import qualified Data.Text as T
import Control.Monad.Random.Class
import Control.Monad.Trans.State.Strict
import Control.Monad.Trans.Except
import qualified Numeric.LinearAlgebra as LA
import System.Environment (getArgs)
import Unsafe.Coerce (unsafeCoerce)

import Qubism.Algebra
import Qubism.CReg
import Qubism.StateVec
import Qubism.QGate
import Qubism.QASM.Parser
import Qubism.QASM.ProgState
import Qubism.QASM.Simulation

-- | hands out the forced draws in order (measureQubit asks for getRandomR (0,1), StateVec.hs:123)
newtype Forced a = Forced (State [Double] a) deriving (Functor, Applicative, Monad)

nextDraw :: Forced Double
nextDraw = Forced $ state $ \ds -> case ds of
  (d:rest) -> (d, rest)
  []       -> (2.0, [])

-- The class methods are polymorphic in the result type; the reference only ever draws Doubles
-- (getRandomR (0, 1 :: Double), StateVec.hs:123), so the forced draw is handed back as it is.
instance MonadRandom Forced where
  getRandomR _  = fmap unsafeCoerce nextDraw
  getRandom     = fmap unsafeCoerce nextDraw
  getRandomRs _ = error "Golden: getRandomRs is not used by the reference"
  getRandoms    = error "Golden: getRandoms is not used by the reference"

runForced :: [Double] -> Forced a -> a
runForced ds (Forced m) = evalState m ds

amps :: StateVec n -> [Complex Double]
amps (UnsafeMkStateVec v) = LA.toList v

showAmps :: [Complex Double] -> String
showAmps zs = unwords [ show re ++ " " ++ show im | (re :+ im) <- zs ]

fromAmps :: [Double] -> LA.Vector (Complex Double)
fromAmps xs = LA.fromList (pair xs) where
  pair (a:b:r) = (a :+ b) : pair r
  pair _       = []

-- | run `k` with the type-level width n
withWidth :: Integer -> (forall n . KnownNat n => Proxy n -> r) -> r
withWidth w k = case someNatVal w of
  Just (SomeNat p) -> k p
  Nothing          -> error "negative width"

gateCase :: [String] -> [Double] -> String
gateCase spec vin = withWidth n $ \(_ :: Proxy n) ->
  let v   = UnsafeMkStateVec (fromAmps vin) :: StateVec n
      fin = finite :: Integer -> Finite n
      u th ph la = unitary th ph la
      out = case kind of
        "U"        -> let [q, th, ph, la] = rest in onJust (fin (round (rd q))) (u (rd th) (rd ph) (rd la)) #> v
        "CU"       -> let [c, t, th, ph, la] = rest
                      in controlled (fin (round (rd c))) (onJust (fin (round (rd t))) (u (rd th) (rd ph) (rd la))) #> v
        "CX"       -> let [c, t] = rest in cnot (fin (round (rd c))) (fin (round (rd t))) #> v
        "COLLAPSE" -> let [q, b] = rest in collapse (fin (round (rd q))) (if (round (rd b) :: Int) == 1 then One else Zero) v
        _          -> error ("unknown gate case " ++ kind)
  in showAmps (amps out)
  where (nS : kind : rest) = spec
        n  = read nS :: Integer
        rd = read :: String -> Double

progCase :: String -> String -> [Double] -> IO String
progCase file src draws = do
  parsed <- parseOpenQASM file (T.pack src)
  case parsed of
    Left err  -> pure ("PARSE-ERROR " ++ show err)
    Right ast -> do
      let r = runForced draws (runExceptT (execStateT (runStmt ast) blankState))
      pure $ case r of
        Left e   -> "RUNTIME-ERROR " ++ show e
        Right ps -> intercalate " ; " $
          [ "SV " ++ T.unpack k ++ " " ++ witnessSV sv (showAmps . amps) | (k, sv) <- Map.toList (stVecs ps) ] ++
          [ "CREG " ++ T.unpack k ++ " " ++ show cr | (k, cr) <- Map.toList (cregs ps) ]

main :: IO ()
main = do
  [inp] <- getArgs
  ls <- lines <$> readFile inp
  mapM_ handle ls
  where
    handle l = case words l of
      ("GATE" : i : rest) ->
        let (spec, vin) = break (== "|") rest
        in putStrLn ("GATE " ++ i ++ " " ++ gateCase spec (map read (drop 1 vin)))
      ("PROG" : name : run : file : ds) -> do
        src <- readFile file
        out <- progCase file src (map read ds)
        putStrLn ("PROG " ++ name ++ " " ++ run ++ " " ++ out)
      _ -> pure ()
